"""Slab-partitioned multi-GPU compression (the reference has no multi-GPU path).

The array is cut into block-aligned slabs along its slowest dimension, one per rank (one process
per GPU).  Because zfp's stream order is block order with the slowest dimension outermost
(src/template/compress.c:72-74), a slab is a contiguous range of blocks AND a contiguous range of
the stream:

* fixed rate: slab g starts at bit  start + blocks_before(g) * maxbits  - no communication;
* variable rate: every rank encodes its slab, the ranks all_gather ONE integer (their slab's bit
  length), an exclusive prefix gives each slab's base bit, and the slab stream is placed there with a
  bit-granular copy (libzfp_b200's zfp_b200_bitcopy on the device).

The per-slab codec is pluggable so that the host logic is testable on CPU (gloo, world_size 2) with
the oracle standing in for the CUDA backend; the default codec is the CUDA backend.
"""
from dataclasses import dataclass

import numpy as np


@dataclass
class SlabPlan:
    rank: int
    world: int
    shape: tuple          # global shape, slowest dimension first
    z0: int               # first index of this rank's slab along the slowest dimension
    z1: int               # one past the last
    blocks_before: int    # blocks of all lower-ranked slabs
    blocks: int           # blocks in this slab

    @property
    def slab_shape(self):
        return (self.z1 - self.z0,) + tuple(self.shape[1:])


def plan_slabs(shape, world):
    """Block-aligned partition of the slowest dimension into `world` contiguous slabs."""
    shape = tuple(int(v) for v in shape)
    layers = (shape[0] + 3) // 4
    per_layer = 1
    for n in shape[1:]:
        per_layer *= (n + 3) // 4
    plans = []
    for g in range(world):
        l0, l1 = g * layers // world, (g + 1) * layers // world
        z0, z1 = min(4 * l0, shape[0]), min(4 * l1, shape[0])
        plans.append(SlabPlan(g, world, shape, z0, z1, l0 * per_layer, (l1 - l0) * per_layer))
    return plans


def place_bits(dst_words, dst_bit, src_words, nbits):
    """Host (numpy) version of the bit-granular placement: OR src[0, nbits) into dst at dst_bit."""
    if nbits == 0:
        return
    nwords = (nbits + 63) // 64
    src = np.ascontiguousarray(src_words[:nwords], dtype=np.uint64).copy()
    if nbits % 64:
        src[-1] &= np.uint64((1 << (nbits % 64)) - 1)
    w0, sh = dst_bit // 64, dst_bit % 64
    if sh == 0:
        dst_words[w0:w0 + nwords] |= src
    else:
        dst_words[w0:w0 + nwords] |= src << np.uint64(sh)
        spill = src >> np.uint64(64 - sh)
        hi = dst_words[w0 + 1:w0 + 1 + nwords]
        hi |= spill[:len(hi)]


class CudaCodec:
    """Per-slab codec backed by libzfp_b200 (device tensors)."""

    def compress(self, slab, mode):
        import zfp_b200
        c = zfp_b200.compress(slab, **mode)
        lengths = c.stream.index_lengths() if not zfp_b200.api.is_fixed_rate_mode(mode) else None
        nbits = int(lengths.astype(np.int64).sum()) if lengths is not None else None
        return c, nbits

    def fixed_bits_per_block(self, dtype_name, dims, mode):
        import zfp_b200
        return zfp_b200.api.mode_params(mode, dtype_name, dims)[1]


def slab_base_bits(local_bits, start_bit=0, group=None):
    """all_gather the slab bit lengths and return (base bit of this rank, list of all lengths).

    One small collective (one int64 per rank) - the only communication of the variable-rate path."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mine = torch.tensor([int(local_bits)], dtype=torch.int64, device=dev)
    everyone = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(everyone, mine, group=group)
    lengths = [int(t.item()) for t in everyone]
    return start_bit + sum(lengths[:rank]), lengths


def fixed_rate_base_bit(plan, maxbits, start_bit=0):
    """Deterministic slab offset for fixed-rate streams - no communication."""
    return start_bit + plan.blocks_before * maxbits


# --------------------------------------------------------------------------------------------------
# device path: one process per GPU (torch.distributed, NCCL)
# --------------------------------------------------------------------------------------------------
def compress_slab_cuda(slab, plan, mode, start_bit=0, group=None):
    """Compress this rank's slab (a CUDA tensor, plan.slab_shape) and locate it in the global stream.

    Returns (Compressed of the slab encoded at bit 0 of a local buffer, slab bit length, base bit of the
    slab in the global stream, list of all slab bit lengths or None for fixed rate)."""
    import zfp_b200
    from zfp_b200 import api
    c = zfp_b200.compress(slab, **mode)
    if api.is_fixed_rate_mode(mode):
        maxbits = api.mode_params(mode, str(slab.dtype), slab.dim())[1]
        return c, plan.blocks * maxbits, fixed_rate_base_bit(plan, maxbits, start_bit), None
    nbits = c.stream.index_bits()  # recorded by the encode: no copy of the per-block index to the host
    base, lengths = slab_base_bits(nbits, start_bit, group)
    return c, nbits, base, lengths


def gather_stream_cuda(c, nbits, base, total_bits, group=None):
    """Assemble the global stream on every rank: all_gather the (padded) slab payloads over NVLink and
    place each at its base bit with the device bit-copy kernel.  Returns an int64 CUDA tensor of words."""
    import torch
    import torch.distributed as dist
    from zfp_b200 import api
    world = dist.get_world_size(group)
    dev = c.words.device
    meta = torch.tensor([nbits, base], dtype=torch.int64, device=dev)
    metas = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    metas = [[int(v) for v in m.tolist()] for m in metas]
    max_words = max((m[0] + 63) // 64 for m in metas) + 1
    mine = torch.zeros(max_words, dtype=torch.int64, device=dev)
    mine[: (nbits + 63) // 64] = c.words[: (nbits + 63) // 64]
    parts = [torch.zeros(max_words, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    out = torch.zeros((total_bits + 63) // 64 + 1, dtype=torch.int64, device=dev)
    for (nb, bs), part in zip(metas, parts):
        api.bitcopy(out, bs, part, 0, nb)
    torch.cuda.synchronize()
    return out


# --------------------------------------------------------------------------------------------------
# stream-ordered device path: no host round trip between the encode, the exchange and the placement
# --------------------------------------------------------------------------------------------------
class DeviceSlab:
    """One rank's slab codec with everything the variable-rate exchange needs kept on the device.

    compress() enqueues, on the current CUDA stream: the slab encode (zfp_b200_encode_async: the slab's bit
    length stays in device memory), the all_gather of the slab lengths (NCCL, same stream order) and the
    device-side prefix / placement (zfp_b200_bitcopy_ranked).  Nothing is read back; `lengths` and `base`
    are device tensors the caller may look at whenever it chooses to synchronise."""

    def __init__(self, slab_shape, dtype, mode, rank, world, group=None, device=None):
        import torch
        from . import api
        self.torch, self.api = torch, api
        self.L = api.load_library()
        self.rank, self.world, self.group = rank, world, group
        self.shape, self.dtype, self.mode = tuple(slab_shape), dtype, dict(mode)
        dev = device or torch.device("cuda", torch.cuda.current_device())
        name = str(dtype).split(".")[-1]
        mn, mx, mp, me = api.mode_params(self.mode, name, len(self.shape))
        d = api.Desc()
        d.type, d.dims = api.ZFP_TYPE[name], len(self.shape)
        for i, n in enumerate(reversed(self.shape)):
            d.n[i], d.s[i] = n, 0
        d.minbits, d.maxbits, d.maxprec, d.minexp = mn, mx, mp, me
        self.desc = d
        self.fixed = mn == mx
        self.words = torch.zeros(self.L.zfp_b200_capacity(api.C.byref(d), 0) // 8 + 2, dtype=torch.int64, device=dev)
        self.index = None if self.fixed else self.L.zfp_b200_index_create()
        self.my_bits = torch.zeros(1, dtype=torch.int64, device=dev)
        self.lengths = torch.zeros(world, dtype=torch.int64, device=dev)
        self.base = torch.zeros(1, dtype=torch.int64, device=dev)
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)  # decode-time index check, accumulated over calls

    def close(self):
        if self.index:
            self.L.zfp_b200_index_destroy(self.index)
            self.index = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def compress(self, slab, global_words=None, start_bit=0):
        """slab: contiguous CUDA tensor of self.shape.  Optionally also places the slab stream in
        `global_words` (an int64 CUDA tensor, zeroed where slabs will land)."""
        torch, api = self.torch, self.api
        import torch.distributed as dist
        assert slab.is_contiguous() and tuple(slab.shape) == self.shape
        st = torch.cuda.current_stream().cuda_stream
        rc = self.L.zfp_b200_encode_async(api.C.byref(self.desc), slab.data_ptr(), self.words.data_ptr(), 0,
                                          self.my_bits.data_ptr(), self.index, st)
        if rc:
            raise RuntimeError("zfp_b200_encode_async failed: %s" % api.last_error())
        if self.world > 1:
            dist.all_gather_into_tensor(self.lengths, self.my_bits, group=self.group)  # stream-ordered, no host sync
        else:
            self.lengths.copy_(self.my_bits)
        rc = self.L.zfp_b200_bitcopy_ranked(global_words.data_ptr() if global_words is not None else None, start_bit,
                                            self.lengths.data_ptr(), self.rank, self.words.data_ptr(), self.base.data_ptr(), st)
        if rc:
            raise RuntimeError("zfp_b200_bitcopy_ranked failed: %s" % api.last_error())

    def decompress(self, out):
        """Decode this rank's slab from its local stream (variable rate: through the block index the encode
        left), stream ordered: `status` (device) turns non-zero if a block did not parse to its indexed length."""
        torch, api = self.torch, self.api
        st = torch.cuda.current_stream().cuda_stream
        rc = self.L.zfp_b200_decode_async(api.C.byref(self.desc), out.data_ptr(), self.words.data_ptr(), 0, self.index,
                                          self.status.data_ptr(), st)
        if rc:
            raise RuntimeError("zfp_b200_decode_async failed: %s" % api.last_error())
        return out
