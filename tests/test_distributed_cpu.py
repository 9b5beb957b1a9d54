"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: slab planning, deterministic fixed-rate
offsets, the all_gather of slab bit lengths for variable rate, and bit-granular placement.  The
per-slab codec is the oracle here (allowed in tests); the result must equal the oracle's stream for
the whole array, i.e. what one GPU - or the serial reference - produces."""
import os
import tempfile

import numpy as np
import pytest

from helpers import make_field


def _worker(rank, world, port_no, shape, dtype, mode, start_bit, tmpdir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.distributed as dist
    from oracle.oracle import Port
    from zfp_b200 import distributed as zd
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port_no, rank=rank, world_size=world)
    try:
        P = Port()
        a = make_field(shape, dtype, seed=11, kind="smooth")
        plan = zd.plan_slabs(shape, world)[rank]
        slab = np.ascontiguousarray(a[plan.z0:plan.z1])
        n = list(reversed(slab.shape)) + [0] * (4 - slab.ndim)
        if slab.size:
            words, end = P.compress_raw(slab.reshape(-1), 0, dtype, n, None, mode)
        else:
            words, end = np.zeros(0, dtype=np.uint64), 0
        if "rate" in mode:
            maxbits = P.params(mode, dtype, len(shape))[1]
            assert end == plan.blocks * maxbits
            base = zd.fixed_rate_base_bit(plan, maxbits, start_bit)
            lengths = None
        else:
            base, lengths = zd.slab_base_bits(end, start_bit)
        total_words = 4 + (start_bit + P.maximum_size(mode, dtype, list(reversed(shape)) + [0] * (4 - len(shape))) * 8) // 64
        out = np.zeros(total_words, dtype=np.uint64)
        zd.place_bits(out, base, words, end)
        np.save(os.path.join(tmpdir, "part%d.npy" % rank), out)
        np.save(os.path.join(tmpdir, "meta%d.npy" % rank), np.array([base, end], dtype=np.int64))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shape,dtype", [((23, 10, 9), np.float64), ((8, 13), np.float32), ((40,), np.int32), ((9, 5, 6, 7), np.float64)])
@pytest.mark.parametrize("mode", [{"rate": 8}, {"rate": 5.3}, {"accuracy": 1e-3}, {"precision": 14}, {"reversible": True}])
def test_two_rank_slabs_concatenate_to_the_serial_stream(port, shape, dtype, mode):
    import torch.multiprocessing as mp
    if np.dtype(dtype).kind != "f" and "accuracy" in mode:
        pytest.skip("accuracy mode is for floating-point data")
    world, start_bit = 2, 96
    with tempfile.TemporaryDirectory() as tmp:
        port_no = 29500 + (os.getpid() + hash((shape, str(mode))) % 1000) % 2000
        mp.spawn(_worker, args=(world, port_no, shape, dtype, mode, start_bit, tmp), nprocs=world, join=True)
        parts = [np.load(os.path.join(tmp, "part%d.npy" % r)) for r in range(world)]
        metas = [np.load(os.path.join(tmp, "meta%d.npy" % r)) for r in range(world)]
    merged = parts[0] | parts[1]
    a = make_field(shape, dtype, seed=11, kind="smooth")
    n = list(reversed(shape)) + [0] * (4 - len(shape))
    want, end = port.compress_raw(a.reshape(-1), 0, dtype, n, None, mode, start_bit=start_bit)
    assert int(metas[0][0]) == start_bit and int(metas[1][0]) == start_bit + int(metas[0][1])
    assert int(metas[1][0] + metas[1][1]) == end
    assert merged[:len(want)].tobytes() == want.tobytes()
    assert not merged[len(want):].any()


def test_plan_slabs_is_block_aligned_and_complete():
    from zfp_b200.distributed import plan_slabs
    for shape in [(2048, 2048, 2048), (1030, 7, 9), (5, 100), (3,), (64, 64, 64, 64)]:
        for world in (1, 2, 4, 8):
            plans = plan_slabs(shape, world)
            assert plans[0].z0 == 0 and plans[-1].z1 == shape[0]
            per_layer = int(np.prod([(n + 3) // 4 for n in shape[1:]])) if len(shape) > 1 else 1
            total = 0
            for i, p in enumerate(plans):
                assert p.z0 % 4 == 0 and (p.z1 % 4 == 0 or p.z1 == shape[0])
                if i:
                    assert p.z0 == plans[i - 1].z1 and p.blocks_before == total
                total += p.blocks
            assert total == ((shape[0] + 3) // 4) * per_layer


def test_place_bits_matches_bitwise_reference():
    from zfp_b200.distributed import place_bits
    rng = np.random.default_rng(5)
    for _ in range(200):
        nbits = int(rng.integers(0, 400))
        dst_bit = int(rng.integers(0, 300))
        src = rng.integers(0, 2 ** 63, size=8, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=8, dtype=np.uint64)
        dst = np.zeros(16, dtype=np.uint64)
        place_bits(dst, dst_bit, src, nbits)
        want = np.zeros(16 * 64, dtype=np.uint8)
        bits = np.unpackbits(src.view(np.uint8), bitorder="little")[:nbits]
        want[dst_bit:dst_bit + nbits] = bits
        assert np.packbits(want, bitorder="little").view(np.uint64).tobytes() == dst.tobytes()
