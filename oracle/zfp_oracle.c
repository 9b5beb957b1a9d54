/*
 * zfp_oracle.c - CPU restatement of zfp's whole-array compress/decompress path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the sm_100a backend in
 * zfp_b200/csrc.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load it.  The product library (libzfp_b200.so) never links, loads or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement against
 *   (1) the reference's own golden checksum tables (tests/constants/checksums in the
 *       reference, extracted into tests/golden/ref_checksums.json) on the reference's seeded
 *       smooth fields, all 4 types x 1-4 D x {rate, precision, accuracy, reversible};
 *   (2) the unmodified reference library built from /root/reference (oracle/_ref), byte for
 *       byte, on seeded inputs including partial blocks, strides, IEEE special values;
 *   (3) committed known-answer fixtures in tests/golden/kat.json.
 *
 * The algorithm is restated in "plane string" form rather than transliterated: every block is
 * turned into a list of unsigned coefficients, and the embedded coder writes, for each bit
 * plane from the top, the bits of already-significant coefficients verbatim followed by a
 * unary group-tested run-length code for the rest; the budgeted and unbudgeted variants of the
 * reference are the same string, truncated at the bit budget.  Each function cites the
 * reference code (path:line under /root/reference) whose behaviour it reproduces.
 *
 * Everything is written for the reference's default build: 64-bit stream words,
 * ZFP_ROUND_NEVER, no DAZ, no tight-error (reference CMakeLists.txt:119-143).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "../zfp_b200/csrc/zfp_perm_tables.h"

#define ZO_MIN_EXP (-1074)   /* include/zfp.h:21 */
#define ZO_MAX_BITS 16658    /* include/zfp.h:19 */
#define ZO_MAX_PREC 64       /* include/zfp.h:20 */

enum { ZO_INT32 = 1, ZO_INT64 = 2, ZO_FLOAT = 3, ZO_DOUBLE = 4 }; /* include/zfp.h:122-128 */

typedef struct {
  int type;        /* zfp_type numbering */
  int dims;        /* 1..4 */
  size_t n[4];     /* nx, ny, nz, nw (unused dims may be 0) */
  ptrdiff_t s[4];  /* element strides; 0 means "default contiguous" (src/template/compress.c:66-68) */
  unsigned minbits, maxbits, maxprec;
  int minexp;
} zo_params;

static const unsigned char perm1[4] = ZFP_B200_PERM1_INIT;
static const unsigned char perm2[16] = ZFP_B200_PERM2_INIT;
static const unsigned char perm3[64] = ZFP_B200_PERM3_INIT;
static const unsigned char perm4[256] = ZFP_B200_PERM4_INIT;
static const unsigned char* const perms[5] = { 0, perm1, perm2, perm3, perm4 };

/* ------------------------------------------------------------------------------------------
 * bit I/O on 64-bit words, least significant bit first (include/zfp/bitstream.inl:240-313).
 * The writer ORs into a buffer the caller has zeroed beyond the start offset, which equals the
 * reference's buffered writes followed by stream_flush's zero padding (bitstream.inl:402-409).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  uint64_t* w;
  uint64_t pos;    /* next bit to write */
  uint64_t limit;  /* bits at or beyond this offset are dropped (budget truncation) */
} sink;

static void put_bit(sink* s, unsigned bit)
{
  if (s->pos < s->limit && bit)
    s->w[s->pos >> 6] |= (uint64_t)1 << (s->pos & 63);
  s->pos++;
}

static void put_bits(sink* s, uint64_t v, unsigned n)
{
  unsigned i;
  for (i = 0; i < n; i++)
    put_bit(s, (unsigned)((v >> i) & 1u));
}

typedef struct {
  const uint64_t* w;
  uint64_t pos;
} source;

static unsigned get_bit(source* s)
{
  unsigned b = (unsigned)((s->w[s->pos >> 6] >> (s->pos & 63)) & 1u);
  s->pos++;
  return b;
}

static uint64_t get_bits(source* s, unsigned n)
{
  uint64_t v = 0;
  unsigned i;
  for (i = 0; i < n; i++)
    v |= (uint64_t)get_bit(s) << i;
  return v;
}

/* ------------------------------------------------------------------------------------------
 * integer helpers: all arithmetic wraps at the width of the scalar's integer type
 * (int32 for float/int32, int64 for double/int64), as the reference's Int/UInt do.
 * ---------------------------------------------------------------------------------------- */
static int64_t wrap(uint64_t v, int P) { return P == 32 ? (int64_t)(int32_t)(uint32_t)v : (int64_t)v; }
static int64_t add(int64_t a, int64_t b, int P) { return wrap((uint64_t)a + (uint64_t)b, P); }
static int64_t sub(int64_t a, int64_t b, int P) { return wrap((uint64_t)a - (uint64_t)b, P); }
static int64_t asr1(int64_t a) { return a >> 1; } /* arithmetic shift, as gcc does for signed */

/* forward lifting of one 4-vector (src/template/encode.c:30-56): two rounds of
 * (average, difference) pairs followed by the fractional rotation of the two odd outputs */
static void lift_fwd(int64_t* p, ptrdiff_t s, int P)
{
  int64_t x = p[0], y = p[s], z = p[2 * s], w = p[3 * s];
  x = asr1(add(x, w, P)); w = sub(w, x, P);
  z = asr1(add(z, y, P)); y = sub(y, z, P);
  x = asr1(add(x, z, P)); z = sub(z, x, P);
  w = asr1(add(w, y, P)); y = sub(y, w, P);
  w = add(w, asr1(y), P); y = sub(y, asr1(w), P);
  p[0] = x; p[s] = y; p[2 * s] = z; p[3 * s] = w;
}

/* inverse lifting (src/template/decode.c:8-45), the UB-free "b += a; a = 2a - b" form */
static void lift_inv(int64_t* p, ptrdiff_t s, int P)
{
  int64_t x = p[0], y = p[s], z = p[2 * s], w = p[3 * s];
  y = add(y, asr1(w), P); w = sub(w, asr1(y), P);
  y = add(y, w, P); w = sub(w, sub(y, w, P), P);
  z = add(z, x, P); x = sub(x, sub(z, x, P), P);
  y = add(y, z, P); z = sub(z, sub(y, z, P), P);
  w = add(w, x, P); x = sub(x, sub(w, x, P), P);
  p[0] = x; p[s] = y; p[2 * s] = z; p[3 * s] = w;
}

/* reversible transform: iterated differences / running sums (src/template/revencode.c:6-38,
 * revdecode.c:6-38) */
static void rlift_fwd(int64_t* p, ptrdiff_t s, int P)
{
  int64_t x = p[0], y = p[s], z = p[2 * s], w = p[3 * s];
  w = sub(w, z, P); z = sub(z, y, P); y = sub(y, x, P);
  w = sub(w, z, P); z = sub(z, y, P);
  w = sub(w, z, P);
  p[0] = x; p[s] = y; p[2 * s] = z; p[3 * s] = w;
}

static void rlift_inv(int64_t* p, ptrdiff_t s, int P)
{
  int64_t x = p[0], y = p[s], z = p[2 * s], w = p[3 * s];
  w = add(w, z, P);
  z = add(z, y, P); w = add(w, z, P);
  y = add(y, x, P); z = add(z, y, P); w = add(w, z, P);
  p[0] = x; p[s] = y; p[2 * s] = z; p[3 * s] = w;
}

typedef void (*lift_fn)(int64_t*, ptrdiff_t, int);

/* apply a 4-point transform along one axis of a 4^dims block */
static void along_axis(int64_t* blk, int dims, int axis, lift_fn f, int P)
{
  int size = 1 << (2 * dims), stride = 1 << (2 * axis), i;
  for (i = 0; i < size; i++)
    if (((i >> (2 * axis)) & 3) == 0)
      f(blk + i, stride, P);
}

/* axis order x,y,z,w forward (encode{1..4}.c fwd_xform), reversed for the inverse
 * (decode{1..4}.c inv_xform); the 1-D lifts along one axis are independent so the order of
 * lines within an axis does not matter */
static void xform_fwd(int64_t* blk, int dims, lift_fn f, int P)
{
  int a;
  for (a = 0; a < dims; a++)
    along_axis(blk, dims, a, f, P);
}

static void xform_inv(int64_t* blk, int dims, lift_fn f, int P)
{
  int a;
  for (a = dims - 1; a >= 0; a--)
    along_axis(blk, dims, a, f, P);
}

/* two's complement <-> negabinary (encode.c:75-80, decode.c:63-68) */
static uint64_t to_negabinary(int64_t x, int P)
{
  uint64_t m = P == 32 ? 0xaaaaaaaaull : 0xaaaaaaaaaaaaaaaaull;
  uint64_t u = (((uint64_t)x + m) ^ m);
  return P == 32 ? (u & 0xffffffffull) : u;
}

static int64_t from_negabinary(uint64_t u, int P)
{
  uint64_t m = P == 32 ? 0xaaaaaaaaull : 0xaaaaaaaaaaaaaaaaull;
  return wrap((u ^ m) - m, P);
}

/* ------------------------------------------------------------------------------------------
 * embedded coder (src/template/encode.c:91-256, decode.c:79-278)
 * ---------------------------------------------------------------------------------------- */

/* Emit the plane strings of `size` coefficients for planes intprec-1 .. kmin into `out`,
 * truncated at `budget` bits.  Returns the number of bits used (<= budget).  Covers
 * encode_few_ints, encode_many_ints and both *_prec variants: they differ only in whether the
 * budget can bind (src/template/codec.c:2-6). */
/* optional coder statistics for kernel design work (planes visited, group-test runs, clusters of
 * adjacent one-bits, verbatim bits); not part of the parity surface */
uint64_t zo_stats[8];

static unsigned encode_ints(sink* out, unsigned budget, unsigned maxprec, const uint64_t* u, unsigned size, unsigned intprec)
{
  unsigned kmin = intprec > maxprec ? intprec - maxprec : 0;
  uint64_t start = out->pos, saved_limit = out->limit;
  unsigned n = 0, k, i;
  uint64_t used;

  if (start + budget < out->limit)
    out->limit = start + budget;
  for (k = intprec; k-- > kmin && out->pos - start < budget;) {
    /* bits of the n coefficients already known to be significant, verbatim */
    zo_stats[0]++;
    zo_stats[3] += n;
    { unsigned j, any = 0; for (j = n; j < size; j++) any |= (unsigned)((u[j] >> k) & 1u); zo_stats[6] += any; }
    for (i = 0; i < n; i++)
      put_bit(out, (unsigned)((u[i] >> k) & 1u));
    /* the rest of the plane: "is there another one-bit?" then its distance in unary */
    while (n < size) {
      unsigned next = n;
      while (next < size && !((u[next] >> k) & 1u))
        next++;
      put_bit(out, next < size);
      if (next == size)
        break;
      zo_stats[1]++;
      if (next != n || n == 0 || !((u[n - 1] >> k) & 1u) ) {
        zo_stats[2]++; /* a run that does not directly continue a previous one-bit starts a cluster */
        if (next == n) zo_stats[4]++; /* ... with no zeros in front of it */
        zo_stats[5] += next - n;
      }
      for (; n < next; n++)
        put_bit(out, 0);
      if (n < size - 1)
        put_bit(out, 1); /* the one at the last position is implied (encode.c:116) */
      n++;
    }
  }
  used = out->pos - start;
  if (used > budget)
    used = budget;
  out->pos = start + used;
  out->limit = saved_limit;
  return (unsigned)used;
}

/* Mirror of the above.  The one subtle point (decode.c:103-111): after a positive group test
 * the one-bit is deposited where the scan stopped even if the budget ran out first. */
static unsigned decode_ints(source* in, unsigned budget, unsigned maxprec, uint64_t* u, unsigned size, unsigned intprec)
{
  unsigned kmin = intprec > maxprec ? intprec - maxprec : 0;
  unsigned bits = budget, n = 0, k, i;

  for (i = 0; i < size; i++)
    u[i] = 0;
  for (k = intprec; bits && k-- > kmin;) {
    unsigned m = n < bits ? n : bits;
    bits -= m;
    for (i = 0; i < m; i++)
      u[i] |= (uint64_t)get_bit(in) << k;
    while (bits && n < size) {
      bits--;
      if (!get_bit(in))
        break;
      while (bits && n < size - 1) {
        bits--;
        if (get_bit(in))
          break;
        n++;
      }
      u[n] |= (uint64_t)1 << k;
      n++;
    }
  }
  return budget - bits;
}

/* lossy integer block: transform, reorder, code, pad (encode.c:259-280) */
static unsigned encode_int_block(sink* out, unsigned minbits, unsigned maxbits, unsigned maxprec, int64_t* blk, int dims, int P)
{
  uint64_t u[256];
  unsigned size = 1u << (2 * dims), i, bits;
  const unsigned char* perm = perms[dims];
  xform_fwd(blk, dims, lift_fwd, P);
  for (i = 0; i < size; i++)
    u[i] = to_negabinary(blk[perm[i]], P);
  bits = encode_ints(out, maxbits, maxprec, u, size, (unsigned)P);
  if (bits < minbits) {
    out->pos += minbits - bits; /* zero padding (encode.c:274-278) */
    bits = minbits;
  }
  return bits;
}

static unsigned decode_int_block(source* in, unsigned minbits, unsigned maxbits, unsigned maxprec, int64_t* blk, int dims, int P)
{
  uint64_t u[256];
  unsigned size = 1u << (2 * dims), i, bits;
  const unsigned char* perm = perms[dims];
  bits = decode_ints(in, maxbits, maxprec, u, size, (unsigned)P);
  if (bits < minbits) {
    in->pos += minbits - bits; /* decode.c:288-292 */
    bits = minbits;
  }
  for (i = 0; i < size; i++)
    blk[perm[i]] = from_negabinary(u[i], P);
  xform_inv(blk, dims, lift_inv, P);
  return bits;
}

/* reversible integer block (revencode.c:56-79, revdecode.c:40-55) */
static unsigned rev_encode_int_block(sink* out, unsigned minbits, unsigned maxbits, unsigned maxprec, int64_t* blk, int dims, int P)
{
  uint64_t u[256], any = 0;
  unsigned size = 1u << (2 * dims), i, prec, bits, pbits = P == 32 ? 5 : 6;
  const unsigned char* perm = perms[dims];
  xform_fwd(blk, dims, rlift_fwd, P);
  for (i = 0; i < size; i++) {
    u[i] = to_negabinary(blk[perm[i]], P);
    any |= u[i];
  }
  /* precision = width minus the number of trailing zero bits shared by all coefficients
   * (rev_precision, revencode.c:41-58), clamped to [1, maxprec] */
  prec = 0;
  if (any) {
    unsigned tz = 0;
    while (!((any >> tz) & 1u))
      tz++;
    prec = (unsigned)P - tz;
  }
  if (prec > maxprec) prec = maxprec;
  if (prec < 1) prec = 1;
  put_bits(out, prec - 1, pbits);
  bits = pbits + encode_ints(out, maxbits - pbits, prec, u, size, (unsigned)P);
  if (bits < minbits) {
    out->pos += minbits - bits;
    bits = minbits;
  }
  return bits;
}

static unsigned rev_decode_int_block(source* in, unsigned minbits, unsigned maxbits, int64_t* blk, int dims, int P)
{
  uint64_t u[256];
  unsigned size = 1u << (2 * dims), i, pbits = P == 32 ? 5 : 6, bits = pbits;
  unsigned prec = (unsigned)get_bits(in, pbits) + 1;
  const unsigned char* perm = perms[dims];
  bits += decode_ints(in, maxbits - bits, prec, u, size, (unsigned)P);
  if (bits < minbits) {
    in->pos += minbits - bits;
    bits = minbits;
  }
  for (i = 0; i < size; i++)
    blk[perm[i]] = from_negabinary(u[i], P);
  xform_inv(blk, dims, rlift_inv, P);
  return bits;
}

/* ------------------------------------------------------------------------------------------
 * floating-point front end
 * ---------------------------------------------------------------------------------------- */

/* block exponent (encodef.c:10-40).  NaNs never win the "max < f" comparison. */
static int block_exponent(const double* f, unsigned size, int is_float)
{
  int ebias = is_float ? 127 : 1023, e = -ebias;
  double max = 0;
  unsigned i;
  for (i = 0; i < size; i++) {
    double a = fabs(f[i]);
    if (max < a)
      max = a;
  }
  if (max > 0) {
    (void)frexp(max, &e); /* float inputs are exactly representable as double: same exponent */
    if (e < 1 - ebias)
      e = 1 - ebias;
  }
  return e;
}

/* number of bit planes to keep (codecf.c:5-13, default rounding mode) */
static unsigned block_precision(int emax, unsigned maxprec, int minexp, int dims)
{
  int p = emax - minexp + 2 * dims + 2;
  if (p < 0) p = 0;
  return (unsigned)p < maxprec ? (unsigned)p : maxprec;
}

/* (Int)(s * f) with s = 2^(P-2-emax) (encodef.c:43-59).  Computed in the scalar's own
 * precision.  When s overflows to +inf (tiny emax; upstream issue #119) the C cast is
 * undefined; the x86-64 reference build produces the "integer indefinite" value INT_MIN for
 * every out-of-range or NaN product (cvttss2si / cvttsd2si), which we state explicitly. */
static void cast_fwd(int64_t* blk, const double* f, unsigned size, int emax, int is_float)
{
  unsigned i;
  if (is_float) {
    float s = ldexpf(1.0f, 30 - emax);
    for (i = 0; i < size; i++) {
      float p = s * (float)f[i];
      blk[i] = (p >= -2147483648.0f && p < 2147483648.0f) ? (int64_t)(int32_t)p : (int64_t)INT32_MIN;
    }
  }
  else {
    double s = ldexp(1.0, 62 - emax);
    for (i = 0; i < size; i++) {
      double p = s * f[i];
      blk[i] = (p >= -9223372036854775808.0 && p < 9223372036854775808.0) ? (int64_t)p : INT64_MIN;
    }
  }
}

/* (Scalar)i * 2^(emax-(P-2)) (codecf.c:15-32) */
static void cast_inv(const int64_t* blk, double* f, unsigned size, int emax, int is_float)
{
  unsigned i;
  if (is_float) {
    float s = ldexpf(1.0f, emax - 30);
    for (i = 0; i < size; i++)
      f[i] = (double)(s * (float)(int32_t)blk[i]);
  }
  else {
    double s = ldexp(1.0, emax - 62);
    for (i = 0; i < size; i++)
      f[i] = s * (double)blk[i];
  }
}

/* bit pattern helpers for the reversible float path */
static uint64_t fbits(double v, int is_float)
{
  if (is_float) { float g = (float)v; uint32_t b; memcpy(&b, &g, 4); return b; }
  else { uint64_t b; memcpy(&b, &v, 8); return b; }
}

/* lossy floating-point block (encodef.c:62-90) */
static unsigned encode_fp_block(sink* out, const zo_params* z, const double* f, int is_float)
{
  int dims = z->dims, P = is_float ? 32 : 64, ebits = is_float ? 8 : 11, ebias = is_float ? 127 : 1023;
  unsigned size = 1u << (2 * dims), bits = 1;
  int emax = block_exponent(f, size, is_float);
  unsigned maxprec = block_precision(emax, z->maxprec, z->minexp, dims);
  unsigned e = maxprec ? (unsigned)(emax + ebias) : 0;
  if (e) {
    int64_t blk[256];
    bits += (unsigned)ebits;
    put_bits(out, 2 * (uint64_t)e + 1, bits);
    cast_fwd(blk, f, size, emax, is_float);
    bits += encode_int_block(out, z->minbits - (bits < z->minbits ? bits : z->minbits), z->maxbits - bits, maxprec, blk, dims, P);
  }
  else {
    put_bit(out, 0);
    if (z->minbits > bits) {
      out->pos += z->minbits - bits;
      bits = z->minbits;
    }
  }
  return bits;
}

static unsigned decode_fp_block(source* in, const zo_params* z, double* f, int is_float)
{
  int dims = z->dims, P = is_float ? 32 : 64, ebits = is_float ? 8 : 11, ebias = is_float ? 127 : 1023;
  unsigned size = 1u << (2 * dims), bits = 1, i;
  if (get_bit(in)) {
    int64_t blk[256];
    int emax;
    unsigned maxprec;
    bits += (unsigned)ebits;
    emax = (int)get_bits(in, (unsigned)ebits) - ebias;
    maxprec = block_precision(emax, z->maxprec, z->minexp, dims);
    bits += decode_int_block(in, z->minbits - (bits < z->minbits ? bits : z->minbits), z->maxbits - bits, maxprec, blk, dims, P);
    cast_inv(blk, f, size, emax, is_float);
  }
  else {
    for (i = 0; i < size; i++)
      f[i] = 0;
    if (z->minbits > bits) {
      in->pos += z->minbits - bits;
      bits = z->minbits;
    }
  }
  return bits;
}

/* reversible floating-point block (revencodef.c:44-80).  `raw` holds the scalars' bit
 * patterns (needed because the double staging of float values cannot carry signalling NaN
 * payloads through arithmetic; we never do arithmetic on `raw`). */
static unsigned rev_encode_fp_block(sink* out, const zo_params* z, const double* f, const uint64_t* raw, int is_float)
{
  int dims = z->dims, P = is_float ? 32 : 64, ebits = is_float ? 8 : 11, ebias = is_float ? 127 : 1023;
  unsigned size = 1u << (2 * dims), bits = 0, i;
  int64_t blk[256];
  double back[256];
  int emax = block_exponent(f, size, is_float), reversible = 1;
  /* rev_fwd_cast / rev_inv_cast special-case the all-zero exponent (revencodef.c:20-27,
   * revcodecf.c:2-11) */
  if (emax != -ebias) {
    cast_fwd(blk, f, size, emax, is_float);
    cast_inv(blk, back, size, emax, is_float);
  }
  else
    for (i = 0; i < size; i++) { blk[i] = 0; back[i] = 0; }
  for (i = 0; i < size; i++)
    if (fbits(back[i], is_float) != raw[i])
      reversible = 0; /* bitwise compare, so -0, NaN, inf and inexact casts all fail */
  if (reversible) {
    unsigned e = (unsigned)(emax + ebias);
    if (!e) {
      put_bit(out, 0);
      return 1; /* note: no minbits padding on this path (revencodef.c:64-69) */
    }
    put_bits(out, 1, 2);
    put_bits(out, e, (unsigned)ebits);
    bits = 2 + (unsigned)ebits;
  }
  else {
    /* sign-magnitude bit patterns -> two's complement (revencodef.c:29-41) */
    uint64_t tcmask = is_float ? 0x7fffffffull : 0x7fffffffffffffffull;
    for (i = 0; i < size; i++) {
      int64_t x = wrap(raw[i], P);
      blk[i] = x < 0 ? wrap((uint64_t)x ^ tcmask, P) : x;
    }
    put_bits(out, 3, 2);
    bits = 2;
  }
  bits += rev_encode_int_block(out, z->minbits - (bits < z->minbits ? bits : z->minbits), z->maxbits - bits, z->maxprec, blk, dims, P);
  return bits;
}

/* revdecodef.c:22-59; results are returned as bit patterns in `raw` */
static unsigned rev_decode_fp_block(source* in, const zo_params* z, uint64_t* raw, int is_float)
{
  int dims = z->dims, P = is_float ? 32 : 64, ebits = is_float ? 8 : 11, ebias = is_float ? 127 : 1023;
  unsigned size = 1u << (2 * dims), bits = 1, i;
  int64_t blk[256];
  if (get_bit(in)) {
    bits++;
    if (get_bit(in)) {
      uint64_t tcmask = is_float ? 0x7fffffffull : 0x7fffffffffffffffull;
      bits += rev_decode_int_block(in, z->minbits - (bits < z->minbits ? bits : z->minbits), z->maxbits - bits, blk, dims, P);
      for (i = 0; i < size; i++) {
        int64_t x = blk[i];
        if (x < 0)
          x = wrap((uint64_t)x ^ tcmask, P);
        raw[i] = P == 32 ? ((uint64_t)x & 0xffffffffull) : (uint64_t)x;
      }
    }
    else {
      double f[256];
      int emax;
      bits += (unsigned)ebits;
      emax = (int)get_bits(in, (unsigned)ebits) - ebias;
      bits += rev_decode_int_block(in, z->minbits - (bits < z->minbits ? bits : z->minbits), z->maxbits - bits, blk, dims, P);
      if (emax != -ebias)
        cast_inv(blk, f, size, emax, is_float);
      else
        for (i = 0; i < size; i++) f[i] = 0;
      for (i = 0; i < size; i++)
        raw[i] = fbits(f[i], is_float);
    }
  }
  else {
    for (i = 0; i < size; i++)
      raw[i] = 0;
    if (z->minbits > bits) {
      in->pos += z->minbits - bits;
      bits = z->minbits;
    }
  }
  return bits;
}

/* ------------------------------------------------------------------------------------------
 * whole-array drivers (src/template/compress.c:17-109, decompress.c; gather/scatter and
 * partial-block padding encode{1..4}.c, decode{1..4}.c, encode.c:8-27)
 * ---------------------------------------------------------------------------------------- */
static size_t type_bytes(int type) { return (type == ZO_INT32 || type == ZO_FLOAT) ? 4 : 8; }

static uint64_t load_raw(const void* data, ptrdiff_t idx, int type)
{
  if (type_bytes(type) == 4) { uint32_t b; memcpy(&b, (const char*)data + idx * 4, 4); return b; }
  else { uint64_t b; memcpy(&b, (const char*)data + idx * 8, 8); return b; }
}

static void store_raw(void* data, ptrdiff_t idx, int type, uint64_t raw)
{
  if (type_bytes(type) == 4) { uint32_t b = (uint32_t)raw; memcpy((char*)data + idx * 4, &b, 4); }
  else memcpy((char*)data + idx * 8, &raw, 8);
}

/* replicate the pad rule along one axis for a line that has m valid entries (encode.c:8-27):
 * m=1 -> (a,a,a,a), m=2 -> (a,b,b,a), m=3 -> (a,b,c,a); m=0 cannot occur */
static void pad_line(uint64_t* p, unsigned m, ptrdiff_t s)
{
  if (m < 2) p[s] = p[0];
  if (m < 3) p[2 * s] = p[s];
  if (m < 4) p[3 * s] = p[0];
}

typedef struct { ptrdiff_t s[4]; size_t nb[4]; size_t nblocks; } layout;

static void make_layout(const zo_params* z, layout* L)
{
  int d;
  ptrdiff_t contiguous = 1;
  L->nblocks = 1;
  for (d = 0; d < 4; d++) {
    size_t n = d < z->dims ? z->n[d] : 1;
    L->s[d] = (d < z->dims && z->s[d]) ? z->s[d] : contiguous;
    L->nb[d] = (n + 3) / 4;
    L->nblocks *= L->nb[d];
    contiguous *= (ptrdiff_t)n;
  }
}

/* fetch block number b (x fastest) as raw bit patterns, padded; returns nothing */
static void gather_block(const zo_params* z, const layout* L, const void* data, size_t b, uint64_t* raw)
{
  size_t org[4], ext[4];
  unsigned size = 1u << (2 * z->dims), i;
  int d;
  for (d = 0; d < 4; d++) {
    size_t n = d < z->dims ? z->n[d] : 1;
    org[d] = 4 * (b % L->nb[d]);
    b /= L->nb[d];
    ext[d] = n - org[d] < 4 ? n - org[d] : 4;
  }
  for (i = 0; i < size; i++) {
    size_t c[4] = { i & 3u, (i >> 2) & 3u, (i >> 4) & 3u, (i >> 6) & 3u };
    if (c[0] < ext[0] && c[1] < ext[1] && c[2] < ext[2] && c[3] < ext[3]) {
      ptrdiff_t idx = 0;
      for (d = 0; d < z->dims; d++)
        idx += L->s[d] * (ptrdiff_t)(org[d] + c[d]);
      raw[i] = load_raw(data, idx, z->type);
    }
    else
      raw[i] = 0;
  }
  /* pad x lines, then y lines, then z, then w; each pass fills lines whose lower-axis
   * coordinates are already complete, which reproduces the nesting in gather_partial */
  for (d = 0; d < z->dims; d++)
    if (ext[d] < 4)
      for (i = 0; i < size; i++)
        if (((i >> (2 * d)) & 3u) == 0) {
          int ok = 1, h;
          for (h = d + 1; h < z->dims; h++)
            if (((i >> (2 * h)) & 3u) >= ext[h])
              ok = 0;
          if (ok)
            pad_line(raw + i, (unsigned)ext[d], (ptrdiff_t)1 << (2 * d));
        }
}

static void scatter_block(const zo_params* z, const layout* L, void* data, size_t b, const uint64_t* raw)
{
  size_t org[4], ext[4];
  unsigned size = 1u << (2 * z->dims), i;
  int d;
  for (d = 0; d < 4; d++) {
    size_t n = d < z->dims ? z->n[d] : 1;
    org[d] = 4 * (b % L->nb[d]);
    b /= L->nb[d];
    ext[d] = n - org[d] < 4 ? n - org[d] : 4;
  }
  for (i = 0; i < size; i++) {
    size_t c[4] = { i & 3u, (i >> 2) & 3u, (i >> 4) & 3u, (i >> 6) & 3u };
    if (c[0] < ext[0] && c[1] < ext[1] && c[2] < ext[2] && c[3] < ext[3]) {
      ptrdiff_t idx = 0;
      for (d = 0; d < z->dims; d++)
        idx += L->s[d] * (ptrdiff_t)(org[d] + c[d]);
      store_raw(data, idx, z->type, raw[i]);
    }
  }
}

static double raw_to_double(uint64_t raw, int type)
{
  if (type == ZO_FLOAT) { uint32_t b = (uint32_t)raw; float g; memcpy(&g, &b, 4); return (double)g; }
  else { double g; memcpy(&g, &raw, 8); return g; }
}

/* Compress the whole field.  `words` must be zero beyond bit `start`.  Returns the bit offset
 * one past the last block (the caller rounds up to a word for the byte size, as
 * zfp_compress does with stream_flush, src/zfp.c:1116-1119).  If `block_bits` is non-NULL it
 * receives the coded length of every block (the block-offset index the GPU path produces). */
uint64_t zo_compress(const zo_params* z, const void* data, uint64_t* words, uint64_t start, uint16_t* block_bits)
{
  layout L;
  sink out;
  size_t b;
  unsigned size = 1u << (2 * z->dims), i;
  int is_fp = z->type == ZO_FLOAT || z->type == ZO_DOUBLE, is_float = z->type == ZO_FLOAT;
  int P = type_bytes(z->type) == 4 ? 32 : 64;
  int reversible = z->minexp < ZO_MIN_EXP; /* src/template/codec.h:4 */
  make_layout(z, &L);
  out.w = words; out.pos = start; out.limit = ~(uint64_t)0;
  for (b = 0; b < L.nblocks; b++) {
    uint64_t raw[256];
    unsigned bits;
    gather_block(z, &L, data, b, raw);
    if (is_fp) {
      double f[256];
      for (i = 0; i < size; i++)
        f[i] = raw_to_double(raw[i], z->type);
      bits = reversible ? rev_encode_fp_block(&out, z, f, raw, is_float) : encode_fp_block(&out, z, f, is_float);
    }
    else {
      int64_t blk[256];
      for (i = 0; i < size; i++)
        blk[i] = wrap(raw[i], P);
      bits = reversible ? rev_encode_int_block(&out, z->minbits, z->maxbits, z->maxprec, blk, z->dims, P)
                        : encode_int_block(&out, z->minbits, z->maxbits, z->maxprec, blk, z->dims, P);
    }
    if (block_bits)
      block_bits[b] = (uint16_t)bits;
  }
  return out.pos;
}

/* Decompress the whole field; returns the bit offset one past the last block. */
uint64_t zo_decompress(const zo_params* z, void* data, const uint64_t* words, uint64_t start)
{
  layout L;
  source in;
  size_t b;
  unsigned size = 1u << (2 * z->dims), i;
  int is_fp = z->type == ZO_FLOAT || z->type == ZO_DOUBLE, is_float = z->type == ZO_FLOAT;
  int P = type_bytes(z->type) == 4 ? 32 : 64;
  int reversible = z->minexp < ZO_MIN_EXP;
  make_layout(z, &L);
  in.w = words; in.pos = start;
  for (b = 0; b < L.nblocks; b++) {
    uint64_t raw[256];
    if (is_fp) {
      if (reversible)
        rev_decode_fp_block(&in, z, raw, is_float);
      else {
        double f[256];
        decode_fp_block(&in, z, f, is_float);
        for (i = 0; i < size; i++)
          raw[i] = fbits(f[i], is_float);
      }
    }
    else {
      int64_t blk[256];
      if (reversible)
        rev_decode_int_block(&in, z->minbits, z->maxbits, blk, z->dims, P);
      else
        decode_int_block(&in, z->minbits, z->maxbits, z->maxprec, blk, z->dims, P);
      for (i = 0; i < size; i++)
        raw[i] = P == 32 ? ((uint64_t)blk[i] & 0xffffffffull) : (uint64_t)blk[i];
    }
    scatter_block(z, &L, data, b, raw);
  }
  return in.pos;
}

/* ------------------------------------------------------------------------------------------
 * host-side parameter logic (src/zfp.c:711-811), restated for cross-checking the product's
 * host API without a GPU
 * ---------------------------------------------------------------------------------------- */
void zo_set_rate(zo_params* z, double rate, int align)
{
  unsigned n = 1u << (2 * z->dims);
  unsigned bits = (unsigned)floor(n * rate + 0.5);
  if (z->type == ZO_FLOAT && bits < 9) bits = 9;
  if (z->type == ZO_DOUBLE && bits < 12) bits = 12;
  if (align) bits = (bits + 63u) & ~63u;
  z->minbits = z->maxbits = bits;
  z->maxprec = ZO_MAX_PREC;
  z->minexp = ZO_MIN_EXP;
}

void zo_set_precision(zo_params* z, unsigned precision)
{
  z->minbits = 1; z->maxbits = ZO_MAX_BITS;
  z->maxprec = precision ? (precision < ZO_MAX_PREC ? precision : ZO_MAX_PREC) : ZO_MAX_PREC;
  z->minexp = ZO_MIN_EXP;
}

void zo_set_accuracy(zo_params* z, double tolerance)
{
  int emin = ZO_MIN_EXP;
  if (tolerance > 0) { (void)frexp(tolerance, &emin); emin--; }
  z->minbits = 1; z->maxbits = ZO_MAX_BITS; z->maxprec = ZO_MAX_PREC; z->minexp = emin;
}

void zo_set_reversible(zo_params* z)
{
  z->minbits = 1; z->maxbits = ZO_MAX_BITS; z->maxprec = ZO_MAX_PREC; z->minexp = ZO_MIN_EXP - 1;
}

size_t zo_maximum_size(const zo_params* z)
{
  layout L;
  int reversible = z->minexp < ZO_MIN_EXP;
  unsigned values = 1u << (2 * z->dims), maxbits = 0, prec = (unsigned)(8 * type_bytes(z->type));
  make_layout(z, &L);
  switch (z->type) {
    case ZO_INT32:  maxbits = reversible ? 5 : 0; break;
    case ZO_INT64:  maxbits = reversible ? 6 : 0; break;
    case ZO_FLOAT:  maxbits = reversible ? 15 : 9; break;
    case ZO_DOUBLE: maxbits = reversible ? 19 : 12; break;
    default: return 0;
  }
  maxbits += values - 1 + values * (z->maxprec < prec ? z->maxprec : prec);
  if (maxbits > z->maxbits) maxbits = z->maxbits;
  if (maxbits < z->minbits) maxbits = z->minbits;
  return (size_t)(((148 + (uint64_t)L.nblocks * maxbits + 63) & ~(uint64_t)63) / 8);
}

/* ------------------------------------------------------------------------------------------
 * Jenkins one-at-a-time hash as used by the reference's golden tables
 * (tests/utils/zfpHash.c:5-51): 64-bit streams hash low and high halves separately.
 * ---------------------------------------------------------------------------------------- */
static uint32_t oaat_step(uint32_t h, uint32_t v) { h += v; h += h << 10; h ^= h >> 6; return h; }
static uint32_t oaat_done(uint32_t h) { h += h << 3; h ^= h >> 11; h += h << 15; return h; }

uint64_t zo_hash_words64(const uint64_t* w, size_t count)
{
  uint32_t lo = 0, hi = 0;
  size_t i;
  for (i = 0; i < count; i++) {
    lo = oaat_step(lo, (uint32_t)w[i]);
    hi = oaat_step(hi, (uint32_t)(w[i] >> 32));
  }
  return (uint64_t)oaat_done(lo) + ((uint64_t)oaat_done(hi) << 32);
}

uint32_t zo_hash_words32(const uint32_t* w, size_t count)
{
  uint32_t h = 0;
  size_t i;
  for (i = 0; i < count; i++)
    h = oaat_step(h, w[i]);
  return oaat_done(h);
}
