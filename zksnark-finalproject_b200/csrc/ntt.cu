// Radix-2 NTT over BLS12-381 Fr for sm_100a -- replaces
// ark_poly::Radix2EvaluationDomain<Fr>::{fft_in_place, ifft_in_place, get_coset}
// (ark-poly ^0.4.2, /root/reference/Cargo.toml:43), the seven transforms inside
// LibsnarkReduction::witness_map_from_matrices.
//
// Design (not a translation of ark-poly's serial/rayon butterflies):
//   * A transform is a decimation-in-frequency (natural -> bit-reversed) or
//     decimation-in-time (bit-reversed -> natural) network whose log n stages
//     are grouped into PASSES over HBM.  One pass = one kernel: a CTA stages a
//     tile of 2^k x C elements in shared memory (k consecutive index bits
//     [t, t+k), C adjacent columns so that global accesses are contiguous),
//     runs k stages on it and writes it back.  Inside a pass every thread keeps
//     8 elements in registers and does 3 stages (12 butterflies) between two
//     shared-memory exchanges.
//   * iNTT followed by a coset NTT needs NO bit-reversal and no separate
//     coset scaling: DIF leaves coefficients bit-reversed, DIT consumes them
//     that way, the coset powers g^i are folded into the per-level twiddle
//     tables (level with butterfly size m uses g^(n/m) * w_m^j), and the two
//     low-bit passes run fused in one kernel on the same tile.
//   * Twiddles come from per-level tables in HBM/L2 (a 32-byte load is ~100x
//     cheaper than the 255-bit multiplication that would recompute it).
//   * n^-1 factors are folded into the witness map's pointwise constants.
#include <cstring>

#include "common.hpp"

namespace b2z {

namespace {

constexpr uint32_t kMaxTileLog = 11;   // 2048 elements = 64 KiB of data per CTA

__device__ __forceinline__ uint32_t sm_phys(uint32_t i) { return i + (i >> 3); }

// Shared-memory tile: two planes of 16-byte half-elements, padded 1/8 so that a
// quarter-warp touching elements at stride 1, 2, 4 or 8 hits 8 distinct 16-byte
// bank groups.
struct Tile {
  uint4* lo;
  uint4* hi;
  __device__ __forceinline__ FrEl load(uint32_t i) const {
    const uint32_t p = sm_phys(i);
    const uint4 a = lo[p], b = hi[p];
    FrEl r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
  }
  __device__ __forceinline__ void store(uint32_t i, const FrEl& v) const {
    const uint32_t p = sm_phys(i);
    lo[p] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    hi[p] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
  }
};

__device__ __forceinline__ FrEl ldg_fr(const FrEl* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  const uint4 a = __ldg(q), b = __ldg(q + 1);
  FrEl r;
  r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
  r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
  return r;
}
__device__ __forceinline__ FrEl ld_fr(const FrEl* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  const uint4 a = q[0], b = q[1];
  FrEl r;
  r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
  r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
  return r;
}
__device__ __forceinline__ void st_fr(FrEl* p, const FrEl& v) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
  q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

// Geometry of one pass: LOCAL index bits [t, t+k) are transformed, C = 2^logC adjacent
// low columns ride along for coalescing (logC <= t).
struct PassGeom {
  uint32_t t, k, logC;
};

// Everything one pass kernel needs.  Up to three vectors are transformed by one launch (blockIdx.y);
// the witness map's a, b, c always travel together.
//
// Distribution (point-sharded prover, several GPUs): a rank holds n / 2^wbits elements of every vector.  Its
// LOCAL index li is the global index gi with the wbits "ownership" bits [ob, ob + wbits) removed; those bits
// equal the rank.  A pass never transforms an ownership bit, so the butterflies are local; only the twiddle
// index (low bits of the GLOBAL index) and -- when the layout changes -- the store address depend on it:
//   ob_in   position of the ownership bits of the layout the data is in (kNoDist: single GPU)
//   ob_out  position for the layout it is stored in; the element goes to rank (gi >> ob_out) & (W - 1), at that
//           rank's local index, through peer[vec][rank] (NVLink stores to peer memory: the all-to-all transpose
//           between the column-owned and the row-owned layout is fused into the pass that precedes it)
constexpr uint32_t kNoDist = 0xffu;
constexpr int kMaxVec = 3, kMaxRanks = 8;
enum PassFlags : uint32_t { PF_CANONICAL_OUT = 1, PF_POINTWISE_IN = 2 };
struct PassArgs {
  FrEl* data[kMaxVec];
  const FrEl* tw_a;
  const FrEl* tw_b;
  PassGeom g;
  uint32_t flags;
  uint32_t ob_in, ob_out, wbits, me;
  FrEl k1, k2;                          // PF_POINTWISE_IN: element = k1 * data[0] * data[1] - k2 * data[2]
  FrEl* peer[kMaxVec][kMaxRanks];       // ob_out != kNoDist: output bases per vector and rank
};

__device__ __forceinline__ uint32_t insert_bits(uint32_t x, uint32_t pos, uint32_t width, uint32_t v) {
  return ((x >> pos) << (pos + width)) | (v << pos) | (x & ((1u << pos) - 1));
}
__device__ __forceinline__ uint32_t remove_bits(uint32_t x, uint32_t pos, uint32_t width) {
  return ((x >> (pos + width)) << pos) | (x & ((1u << pos) - 1));
}

// One exchange round: Q stages on local bits [p0, p0+Q) of the tile's k-bit row
// index.  Each thread owns 8 elements: 8 >> Q groups of 2^Q.
template <int Q, bool DIT, bool DIST>
__device__ __forceinline__ void ntt_round(const Tile& tile, const FrEl* __restrict__ tw, const PassArgs& a, uint32_t p0,
                                          uint32_t col_base) {
  constexpr int R = 1 << Q;
  constexpr int G = 8 >> Q;
  const PassGeom g = a.g;
  const uint32_t ngroups = (1u << (g.k + g.logC)) >> Q;
  const uint32_t cmask = (1u << g.logC) - 1;
#pragma unroll 1
  for (int gi = 0; gi < G; gi++) {
    const uint32_t gid = threadIdx.x + gi * blockDim.x;
    if (gid < ngroups) {
      const uint32_t c = gid & cmask;
      const uint32_t rest = gid >> g.logC;
      const uint32_t blow = rest & ((1u << p0) - 1);
      const uint32_t bhigh = rest >> p0;
      const uint32_t row0 = (bhigh << (p0 + Q)) | blow;
      FrEl x[R];
#pragma unroll
      for (int e = 0; e < R; e++) x[e] = tile.load((((row0 | ((uint32_t)e << p0))) << g.logC) | c);
#pragma unroll
      for (int s = 0; s < Q; s++) {
        const int r = DIT ? s : Q - 1 - s;
        const uint32_t lvl = g.t + p0 + r;                 // local bit of this stage
        uint32_t glvl = lvl;                               // its global bit = twiddle level
        if (DIST && lvl >= a.ob_in) glvl += a.wbits;
        const FrEl* twl = tw + ((1u << glvl) - 1);
#pragma unroll
        for (int e = 0; e < R; e++) {
          if (e & (1 << r)) continue;
          uint32_t j = ((blow | ((uint32_t)(e & ((1 << r) - 1)) << p0)) << g.t) | (col_base + c);
          if (DIST && lvl >= a.ob_in) j = insert_bits(j, a.ob_in, a.wbits, a.me);
          const FrEl w = ldg_fr(twl + j);
          FrEl& u = x[e];
          FrEl& v = x[e | (1 << r)];
          if (DIT) {
            const FrEl tv = Fr::mul(w, v);
            v = Fr::sub(u, tv);
            u = Fr::add(u, tv);
          } else {
            const FrEl d = Fr::sub(u, v);
            u = Fr::add(u, v);
            v = Fr::mul(w, d);
          }
        }
      }
#pragma unroll
      for (int e = 0; e < R; e++) tile.store((((row0 | ((uint32_t)e << p0))) << g.logC) | c, x[e]);
    }
  }
  __syncthreads();
}

template <bool DIT, bool DIST>
__device__ __forceinline__ void ntt_round_q(int q, const Tile& tile, const FrEl* tw, const PassArgs& a, uint32_t p0,
                                            uint32_t col_base) {
  if (q == 3) ntt_round<3, DIT, DIST>(tile, tw, a, p0, col_base);
  else if (q == 2) ntt_round<2, DIT, DIST>(tile, tw, a, p0, col_base);
  else ntt_round<1, DIT, DIST>(tile, tw, a, p0, col_base);
}

// All k stages of a pass on the tile in shared memory.
template <bool DIT, bool DIST>
__device__ __forceinline__ void ntt_tile_stages(const Tile& tile, const FrEl* tw, const PassArgs& a, uint32_t col_base) {
  const uint32_t k = a.g.k;
  if (DIT) {
    uint32_t p0 = 0;
    while (p0 < k) {
      const int q = (k - p0 >= 3) ? 3 : (int)(k - p0);
      ntt_round_q<true, DIST>(q, tile, tw, a, p0, col_base);
      p0 += q;
    }
  } else {
    uint32_t top = k;
    while (top > 0) {
      const int q = top >= 3 ? 3 : (int)top;
      ntt_round_q<false, DIST>(q, tile, tw, a, top - q, col_base);
      top -= q;
    }
  }
}

enum PassMode { PASS_DIF = 0, PASS_DIT = 1, PASS_DIF_DIT = 2 };

// grid = (n_local / 2^(k+logC) tiles, vectors); block = max(32, tile/8) threads;
// dynamic smem = 2 * (S + S/8) * 16 bytes.
template <int MODE, bool DIST>
__global__ void __launch_bounds__(256, 2) ntt_pass_kernel(const __grid_constant__ PassArgs a) {
  extern __shared__ uint4 smem[];
  const PassGeom g = a.g;
  const uint32_t S = 1u << (g.k + g.logC);
  Tile tile{smem, smem + S + (S >> 3)};
  const uint32_t ltiles_log = g.t - g.logC;
  const uint32_t tileL = blockIdx.x & ((1u << ltiles_log) - 1);
  const uint64_t H = blockIdx.x >> ltiles_log;
  const uint32_t col_base = tileL << g.logC;
  const uint64_t base_idx = (H << (g.t + g.k)) | col_base;
  FrEl* base = a.data[blockIdx.y] + base_idx;
  const uint32_t cmask = (1u << g.logC) - 1;
  if (a.flags & PF_POINTWISE_IN) {
    const FrEl* pb = a.data[1] + base_idx;
    const FrEl* pc = a.data[2] + base_idx;
    for (uint32_t i = threadIdx.x; i < S; i += blockDim.x) {
      const uint64_t off = ((uint64_t)(i >> g.logC) << g.t) | (i & cmask);
      const FrEl ab = Fr::mul(Fr::reduce(ld_fr(base + off)), ld_fr(pb + off));
      tile.store(i, Fr::sub(Fr::mul(a.k1, ab), Fr::mul(a.k2, ld_fr(pc + off))));
    }
  } else {
    for (uint32_t i = threadIdx.x; i < S; i += blockDim.x) {
      const uint64_t off = ((uint64_t)(i >> g.logC) << g.t) | (i & cmask);
      tile.store(i, ld_fr(base + off));
    }
  }
  __syncthreads();
  if (MODE == PASS_DIF) ntt_tile_stages<false, DIST>(tile, a.tw_a, a, col_base);
  if (MODE == PASS_DIT) ntt_tile_stages<true, DIST>(tile, a.tw_a, a, col_base);
  if (MODE == PASS_DIF_DIT) {
    ntt_tile_stages<false, DIST>(tile, a.tw_a, a, col_base);
    ntt_tile_stages<true, DIST>(tile, a.tw_b, a, col_base);
  }
  const bool canon = (a.flags & PF_CANONICAL_OUT) != 0;
  if (DIST && a.ob_out != kNoDist) {
    // layout change fused into the store: the element's global index decides the rank and the slot
    for (uint32_t i = threadIdx.x; i < S; i += blockDim.x) {
      const uint32_t li = (uint32_t)base_idx + (((i >> g.logC) << g.t) | (i & cmask));
      const uint32_t gi = insert_bits(li, a.ob_in, a.wbits, a.me);
      const uint32_t dst_rank = (gi >> a.ob_out) & ((1u << a.wbits) - 1);
      FrEl v = tile.load(i);
      if (canon) v = Fr::reduce(v);
      st_fr(a.peer[blockIdx.y][dst_rank] + remove_bits(gi, a.ob_out, a.wbits), v);
    }
  } else {
    for (uint32_t i = threadIdx.x; i < S; i += blockDim.x) {
      const uint64_t off = ((uint64_t)(i >> g.logC) << g.t) | (i & cmask);
      FrEl v = tile.load(i);
      if (canon) v = Fr::reduce(v);
      st_fr(base + off, v);
    }
  }
}

template <int MODE, bool DIST>
void launch_pass_t(Ctx* ctx, const PassArgs& a, uint32_t log_n_local, uint32_t nvec, cudaStream_t st) {
  const uint32_t ls = a.g.k + a.g.logC;
  const uint32_t S = 1u << ls;
  const uint32_t threads = S / 8 < 32 ? 32 : S / 8;
  const size_t smem = (size_t)2 * (S + (S >> 3)) * sizeof(uint4);
  // per-device function attribute (a process may hold contexts on several GPUs): set it once per context
  const int slot = MODE * 2 + (DIST ? 1 : 0);
  if (!ctx->ntt_attr_set[slot]) {
    B2Z_CUDA(cudaFuncSetAttribute(ntt_pass_kernel<MODE, DIST>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  2 * ((1 << kMaxTileLog) + (1 << (kMaxTileLog - 3))) * (int)sizeof(uint4)));
    ctx->ntt_attr_set[slot] = true;
  }
  const uint32_t grid = 1u << (log_n_local - ls);
  ProfileScope ps(ctx, PH_NTT_PASS, st, (uint64_t)nvec << log_n_local);
  ntt_pass_kernel<MODE, DIST><<<dim3(grid, nvec), threads, smem, st>>>(a);
  B2Z_LAUNCHED(ctx);
}

template <int MODE>
void launch_pass(Ctx* ctx, const PassArgs& a, uint32_t log_n_local, uint32_t nvec, cudaStream_t st) {
  if (a.ob_in != kNoDist) launch_pass_t<MODE, true>(ctx, a, log_n_local, nvec, st);
  else launch_pass_t<MODE, false>(ctx, a, log_n_local, nvec, st);
}

// Pass plan: bits [0, k0) are the contiguous low pass; the remaining high bits
// are split into passes of at most kMaxTileLog bits, each padded with columns up
// to the tile size.  wbits > 0: the plan of a transform distributed over 2^wbits ranks -- the high passes run
// on the column-owned layout (ownership bits [k0 - wbits, k0)), the low pass on the row-owned layout
// (ownership bits [log_n - wbits, log_n)); geometries are in LOCAL bits.
struct Plan {
  int npass = 0;
  PassGeom pass[4];   // pass[0] = low bits, ascending
  uint32_t k0 = 0;
};

Plan make_plan(uint32_t log_n, uint32_t wbits = 0) {
  Plan p;
  const uint32_t log_local = log_n - wbits;
  // tile = 2^ls elements per CTA (128 or 256 threads): full tiles for big transforms, smaller ones for small
  // transforms so that a 2^16..2^18 NTT still spreads over enough CTAs (148 SMs)
  uint32_t ls = log_local > 8 ? log_local - 8 : 0;
  if (ls < 10) ls = 10;
  if (ls > kMaxTileLog) ls = kMaxTileLog;
  uint32_t k0 = log_n < ls ? log_n : ls;
  if (wbits && log_n - k0 < wbits) k0 = log_n - wbits;     // the row-owned layout needs wbits high bits
  p.k0 = k0;
  p.pass[p.npass++] = PassGeom{0, k0, 0};
  uint32_t t = k0 - wbits;                                  // local position of the first high bit
  uint32_t rem = log_n - k0;
  const uint32_t nhi = (rem + ls - 1) / ls;
  for (uint32_t i = 0; i < nhi; i++) {
    const uint32_t k = (rem + (nhi - i) - 1) / (nhi - i);   // balanced split
    uint32_t logC = ls > k ? ls - k : 0;
    if (logC > t) logC = t;
    if (wbits && logC > k0 - wbits) logC = k0 - wbits;      // columns are bits below the ownership bits
    p.pass[p.npass++] = PassGeom{t, k, logC};
    t += k;
    rem -= k;
  }
  return p;
}

PassArgs pass_args(FrEl* v0, FrEl* v1, FrEl* v2, const FrEl* tw_a, const FrEl* tw_b, PassGeom g) {
  PassArgs a;
  std::memset(&a, 0, sizeof(a));
  a.data[0] = v0; a.data[1] = v1; a.data[2] = v2;
  a.tw_a = tw_a; a.tw_b = tw_b;
  a.g = g;
  a.ob_in = kNoDist; a.ob_out = kNoDist;
  return a;
}

__global__ void bitrev_kernel(FrEl* data, uint32_t log_n, FrEl scale, int has_scale, const FrEl* __restrict__ pw_lo,
                              const FrEl* __restrict__ pw_hi) {
  const uint64_t n = 1ull << log_n;
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t j = log_n == 0 ? 0 : (__brevll(i) >> (64 - log_n));
  if (j < i) return;
  // element stored at i belongs at natural index j and vice versa
  FrEl vi = ld_fr(data + i);
  FrEl vj = ld_fr(data + j);
  if (has_scale) { vi = Fr::mul(scale, vi); vj = Fr::mul(scale, vj); }
  if (pw_lo != nullptr) {
    // vi lands at index j, vj at index i
    FrEl wj = Fr::mul(Fr::reduce(ldg_fr(pw_lo + (j & 1023))), ldg_fr(pw_hi + (j >> 10)));
    FrEl wi = Fr::mul(Fr::reduce(ldg_fr(pw_lo + (i & 1023))), ldg_fr(pw_hi + (i >> 10)));
    vi = Fr::mul(Fr::reduce(wj), vi);
    vj = Fr::mul(Fr::reduce(wi), vj);
  }
  vi = Fr::reduce(vi);
  vj = Fr::reduce(vj);
  st_fr(data + j, vi);
  if (i != j) st_fr(data + i, vj);
}

__global__ void scale_powers_kernel(FrEl* data, uint64_t n, const FrEl* __restrict__ pw_lo,
                                    const FrEl* __restrict__ pw_hi) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const FrEl w = Fr::mul(Fr::reduce(ldg_fr(pw_lo + (i & 1023))), ldg_fr(pw_hi + (i >> 10)));
  st_fr(data + i, Fr::mul(Fr::reduce(w), ld_fr(data + i)));
}

__global__ void pow_table_kernel(FrEl* out, FrEl base, uint32_t count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  st_fr(out + i, Fr::reduce(Fr::pow_u64(base, i)));
}

// out[j] = shift * root^j for j < 2^lvl, canonical
__global__ void twiddle_level_kernel(FrEl* out, FrEl root, FrEl shift, uint32_t count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  st_fr(out + i, Fr::reduce(Fr::mul_safe(Fr::pow_u64(root, i), shift)));
}

__global__ void wm_pointwise_kernel(FrEl* a, const FrEl* __restrict__ b, const FrEl* __restrict__ c, uint64_t n,
                                    FrEl k1, FrEl k2) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const FrEl ab = Fr::mul(Fr::reduce(ld_fr(a + i)), ld_fr(b + i));
  const FrEl t1 = Fr::mul(k1, ab);
  const FrEl t2 = Fr::mul(k2, ld_fr(c + i));
  st_fr(a + i, Fr::sub(t1, t2));
}

__global__ void canonicalize_kernel(FrEl* data, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  st_fr(data + i, Fr::reduce(ld_fr(data + i)));
}

inline uint32_t blocks_for(uint64_t n, uint32_t threads) { return (uint32_t)((n + threads - 1) / threads); }

FrEl fr_from_u64_host(uint64_t v) {
  FrEl x = Fr::zero();
  x.l[0] = (uint32_t)v;
  x.l[1] = (uint32_t)(v >> 32);
  return Fr::reduce(Fr::to_mont(x));
}

FrEl fr_const(uint32_t (*f)(int)) {
  FrEl x;
  for (int i = 0; i < 8; i++) x.l[i] = f(i);
  return x;
}

}  // namespace

// ---------------------------------------------------------------------------
// Host-visible launchers
// ---------------------------------------------------------------------------
void ntt_dif(Ctx* ctx, const FrEl* tw, FrEl* data, uint32_t log_n, cudaStream_t st) {
  if (log_n == 0) return;
  const Plan p = make_plan(log_n);
  for (int i = p.npass - 1; i >= 0; i--)
    launch_pass<PASS_DIF>(ctx, pass_args(data, nullptr, nullptr, tw, nullptr, p.pass[i]), log_n, 1, st);
}

void ntt_dit(Ctx* ctx, const FrEl* tw, FrEl* data, uint32_t log_n, cudaStream_t st) {
  if (log_n == 0) return;
  const Plan p = make_plan(log_n);
  for (int i = 0; i < p.npass; i++)
    launch_pass<PASS_DIT>(ctx, pass_args(data, nullptr, nullptr, tw, nullptr, p.pass[i]), log_n, 1, st);
}

// DIF with tw_a then DIT with tw_b on up to three vectors at once (natural -> natural), low stages fused
static void ntt_dif_dit_n(Ctx* ctx, const FrEl* tw_a, const FrEl* tw_b, FrEl* v0, FrEl* v1, FrEl* v2, uint32_t nvec,
                          uint32_t log_n, cudaStream_t st) {
  if (log_n == 0) return;
  const Plan p = make_plan(log_n);
  for (int i = p.npass - 1; i >= 1; i--)
    launch_pass<PASS_DIF>(ctx, pass_args(v0, v1, v2, tw_a, nullptr, p.pass[i]), log_n, nvec, st);
  launch_pass<PASS_DIF_DIT>(ctx, pass_args(v0, v1, v2, tw_a, tw_b, p.pass[0]), log_n, nvec, st);
  for (int i = 1; i < p.npass; i++)
    launch_pass<PASS_DIT>(ctx, pass_args(v0, v1, v2, tw_b, nullptr, p.pass[i]), log_n, nvec, st);
}

void ntt_dif_dit(Ctx* ctx, const FrEl* tw_a, const FrEl* tw_b, FrEl* data, uint32_t log_n, cudaStream_t st) {
  ntt_dif_dit_n(ctx, tw_a, tw_b, data, nullptr, nullptr, 1, log_n, st);
}

void ntt_bitrev(Ctx* ctx, FrEl* data, uint32_t log_n, const FrEl* scale, const FrEl* pw_lo, const FrEl* pw_hi,
                cudaStream_t st) {
  const uint64_t n = 1ull << log_n;
  FrEl s = scale ? *scale : Fr::one();
  bitrev_kernel<<<blocks_for(n, 256), 256, 0, st>>>(data, log_n, s, scale != nullptr, pw_lo, pw_hi);
  B2Z_LAUNCHED(ctx);
}

void ntt_scale_powers(Ctx* ctx, FrEl* data, uint32_t log_n, const FrEl* pw_lo, const FrEl* pw_hi, cudaStream_t st) {
  const uint64_t n = 1ull << log_n;
  scale_powers_kernel<<<blocks_for(n, 256), 256, 0, st>>>(data, n, pw_lo, pw_hi);
  B2Z_LAUNCHED(ctx);
}

void fr_pow_table(Ctx* ctx, FrEl* out, const FrEl& base, uint32_t count, cudaStream_t st) {
  pow_table_kernel<<<blocks_for(count, 128), 128, 0, st>>>(out, base, count);
  B2Z_LAUNCHED(ctx);
}

void wm_pointwise(Ctx* ctx, FrEl* a, const FrEl* b, const FrEl* c, uint32_t log_n, const FrEl& k1, const FrEl& k2,
                  cudaStream_t st) {
  const uint64_t n = 1ull << log_n;
  ProfileScope ps(ctx, PH_WM_POINTWISE, st, n);
  wm_pointwise_kernel<<<blocks_for(n, 256), 256, 0, st>>>(a, b, c, n, k1, k2);
  B2Z_LAUNCHED(ctx);
}

void fr_canonicalize(Ctx* ctx, FrEl* data, size_t n, cudaStream_t st) {
  if (n == 0) return;
  canonicalize_kernel<<<blocks_for(n, 256), 256, 0, st>>>(data, n);
  B2Z_LAUNCHED(ctx);
}

// ---------------------------------------------------------------------------
// Domains and twiddle tables
// ---------------------------------------------------------------------------
const NttDomain& ntt_domain(Ctx* ctx, uint32_t log_n) {
  auto it = ctx->domains.find(log_n);
  if (it != ctx->domains.end()) return it->second;
  NttDomain& d = ctx->domains[log_n];
  d.log_n = log_n;
  const FrEl n_mont = fr_from_u64_host(1ull << log_n);
  d.n_inv = Fr::reduce(Fr::inv(n_mont));
  // Z_H(g) = g^n - 1 on the coset g = 7  (evaluate_vanishing_polynomial, A.2)
  const FrEl g = fr_const(FrParams::generator);
  FrEl gn = g;
  for (uint32_t i = 0; i < log_n; i++) gn = Fr::sqr(gn);
  const FrEl z = Fr::reduce(Fr::sub(gn, Fr::one()));
  const FrEl zinv = Fr::reduce(Fr::inv(z));
  const FrEl n2 = Fr::reduce(Fr::sqr(d.n_inv));
  d.wm_k2 = Fr::reduce(Fr::mul(n2, zinv));
  d.wm_k1 = Fr::reduce(Fr::mul(d.n_inv, d.wm_k2));
  return d;
}

const FrEl* ntt_twiddles(Ctx* ctx, uint32_t log_n, TwKind kind, cudaStream_t st) {
  NttDomain& d = const_cast<NttDomain&>(ntt_domain(ctx, log_n));
  if (d.tw[kind].p != nullptr || log_n == 0) return d.tw[kind].p;
  const uint64_t n = 1ull << log_n;
  d.tw[kind].alloc(n);   // n - 1 used
  const bool inverse = (kind == TW_INV || kind == TW_COSET_INV);
  const bool coset = (kind == TW_COSET_FWD || kind == TW_COSET_INV);
  const FrEl root32 = fr_const(FrParams::root_of_unity);
  FrEl g = fr_const(FrParams::generator);
  if (inverse) g = Fr::reduce(Fr::inv(g));
  for (uint32_t lvl = 0; lvl < log_n; lvl++) {
    // butterfly size m = 2^(lvl+1): root = w_m^(+-1), shift = g^(+-n/m)
    FrEl root = root32;
    for (uint32_t i = 0; i < 32 - (lvl + 1); i++) root = Fr::sqr(root);
    root = Fr::reduce(root);
    if (inverse) root = Fr::reduce(Fr::inv(root));
    FrEl shift = Fr::one();
    if (coset) {
      shift = g;
      for (uint32_t i = 0; i < log_n - (lvl + 1); i++) shift = Fr::sqr(shift);
    }
    shift = Fr::reduce(shift);
    const uint32_t count = 1u << lvl;
    twiddle_level_kernel<<<blocks_for(count, 128), 128, 0, st>>>(d.tw[kind].p + (count - 1), root, shift, count);
    B2Z_LAUNCHED(ctx);
  }
  return d.tw[kind].p;
}

// ---------------------------------------------------------------------------
// Witness map (A.3): h = iNTT_coset( (NTT_coset(iNTT a) * NTT_coset(iNTT b)
//                                     - NTT_coset(iNTT c)) / Z_H(g) )
// All n^-1 factors and Z^-1 are folded into wm_k1, wm_k2.
// ---------------------------------------------------------------------------
// One input vector: evaluations on the domain -> evaluations on the coset g H (unscaled inverse
// transform fused with the coset transform; the n^-1 factors are folded into the pointwise constants).
void witness_map_transform(Ctx* ctx, FrEl* x, uint32_t log_n, cudaStream_t st) {
  const FrEl* tw_inv = ntt_twiddles(ctx, log_n, TW_INV, st);
  const FrEl* tw_cf = ntt_twiddles(ctx, log_n, TW_COSET_FWD, st);
  ntt_dif_dit(ctx, tw_inv, tw_cf, x, log_n, st);
}
// the three input vectors in the same launches (3 launches at 2^17..2^22 instead of 9)
void witness_map_transform3(Ctx* ctx, FrEl* a, FrEl* b, FrEl* c, uint32_t log_n, cudaStream_t st) {
  const FrEl* tw_inv = ntt_twiddles(ctx, log_n, TW_INV, st);
  const FrEl* tw_cf = ntt_twiddles(ctx, log_n, TW_COSET_FWD, st);
  ntt_dif_dit_n(ctx, tw_inv, tw_cf, a, b, c, 3, log_n, st);
}

// a <- coefficients of (a b - c) / Z_H from the three coset evaluation vectors.  The pointwise step is folded
// into the load of the first transform pass, the canonical form into the store of the last one.
void witness_map_quotient(Ctx* ctx, FrEl* a, const FrEl* b, const FrEl* c, uint32_t log_n, bool natural_out,
                          cudaStream_t st) {
  const NttDomain& d = ntt_domain(ctx, log_n);
  const FrEl* tw_ci = ntt_twiddles(ctx, log_n, TW_COSET_INV, st);
  if (log_n == 0) {
    wm_pointwise(ctx, a, b, c, log_n, d.wm_k1, d.wm_k2, st);
    fr_canonicalize(ctx, a, 1, st);
    return;
  }
  const Plan p = make_plan(log_n);
  for (int i = p.npass - 1; i >= 0; i--) {
    PassArgs pa = pass_args(a, const_cast<FrEl*>(b), const_cast<FrEl*>(c), tw_ci, nullptr, p.pass[i]);
    if (i == p.npass - 1) { pa.flags |= PF_POINTWISE_IN; pa.k1 = d.wm_k1; pa.k2 = d.wm_k2; }
    if (i == 0 && !natural_out) pa.flags |= PF_CANONICAL_OUT;
    launch_pass<PASS_DIF>(ctx, pa, log_n, 1, st);
  }
  if (natural_out) ntt_bitrev(ctx, a, log_n, nullptr, nullptr, nullptr, st);
}

void witness_map_device(Ctx* ctx, FrEl* a, FrEl* b, FrEl* c, uint32_t log_n, bool natural_out, cudaStream_t st) {
  witness_map_transform3(ctx, a, b, c, log_n, st);
  witness_map_quotient(ctx, a, b, c, log_n, natural_out, st);
}

// ---------------------------------------------------------------------------
// Distributed witness map (point-sharded prover over W = 2^wbits GPUs).  Every rank holds n / W elements
// of each vector.  Column-owned layout X: ownership bits [k0 - wbits, k0) (the rank owns 2^(k0 - wbits)
// columns of the 2^k0 x 2^(log_n - k0) matrix) -- the high passes are local.  Row-owned layout Y: ownership
// bits [log_n - wbits, log_n) (a contiguous chunk) -- the low pass is local.  A layout change is the store
// of the pass before it, straight into the peers' buffers over NVLink; the caller puts a barrier between a
// step that writes peers and the step that reads what was written.
// ---------------------------------------------------------------------------
bool ntt_dist_supported(uint32_t log_n, uint32_t wbits) {
  if (wbits == 0 || wbits > 3 || log_n < 14 || log_n > 28) return false;
  const Plan p = make_plan(log_n, wbits);
  return p.k0 > wbits && log_n - p.k0 >= wbits;
}
uint32_t ntt_dist_col_bits(uint32_t log_n, uint32_t wbits) { return make_plan(log_n, wbits).k0 - wbits; }

static void dist_fill(PassArgs& pa, const NttDist& D, uint32_t ob_in, uint32_t ob_out, FrEl* const* out_of_rank0,
                      FrEl* const* out_of_rank1, FrEl* const* out_of_rank2) {
  pa.ob_in = ob_in; pa.ob_out = ob_out; pa.wbits = D.wbits; pa.me = D.me;
  if (ob_out != kNoDist)
    for (uint32_t r = 0; r < (1u << D.wbits); r++) {
      pa.peer[0][r] = out_of_rank0 ? out_of_rank0[r] : nullptr;
      pa.peer[1][r] = out_of_rank1 ? out_of_rank1[r] : nullptr;
      pa.peer[2][r] = out_of_rank2 ? out_of_rank2[r] : nullptr;
    }
}

// step 1 (after the row evaluation into X): inverse-transform high passes on X_a, X_b, X_c; the last one
// scatters into the peers' Y_a, Y_b, Y_c
void wm_dist_step1(Ctx* ctx, const NttDist& D, uint32_t log_n, cudaStream_t st) {
  const Plan p = make_plan(log_n, D.wbits);
  const uint32_t obC = p.k0 - D.wbits, obR = log_n - D.wbits, ll = log_n - D.wbits;
  const FrEl* tw_inv = ntt_twiddles(ctx, log_n, TW_INV, st);
  for (int i = p.npass - 1; i >= 1; i--) {
    PassArgs pa = pass_args(D.x[0], D.x[1], D.x[2], tw_inv, nullptr, p.pass[i]);
    dist_fill(pa, D, obC, i == 1 ? obR : kNoDist, D.peer_y[0], D.peer_y[1], D.peer_y[2]);
    launch_pass<PASS_DIF>(ctx, pa, ll, 3, st);
  }
}
// step 2: fused low pass (inverse then coset-forward) on Y, scattering back into the peers' X
void wm_dist_step2(Ctx* ctx, const NttDist& D, uint32_t log_n, cudaStream_t st) {
  const Plan p = make_plan(log_n, D.wbits);
  const uint32_t obC = p.k0 - D.wbits, obR = log_n - D.wbits, ll = log_n - D.wbits;
  PassArgs pa = pass_args(D.y[0], D.y[1], D.y[2], ntt_twiddles(ctx, log_n, TW_INV, st),
                          ntt_twiddles(ctx, log_n, TW_COSET_FWD, st), p.pass[0]);
  dist_fill(pa, D, obR, obC, D.peer_x[0], D.peer_x[1], D.peer_x[2]);
  launch_pass<PASS_DIF_DIT>(ctx, pa, ll, 3, st);
}
// step 3: coset-forward high passes on X (local), then the quotient (pointwise folded into the load) and the
// coset-inverse high passes on X_a; the last one scatters into the peers' Y_a
void wm_dist_step3(Ctx* ctx, const NttDist& D, uint32_t log_n, cudaStream_t st) {
  const Plan p = make_plan(log_n, D.wbits);
  const uint32_t obC = p.k0 - D.wbits, obR = log_n - D.wbits, ll = log_n - D.wbits;
  const NttDomain& d = ntt_domain(ctx, log_n);
  const FrEl* tw_cf = ntt_twiddles(ctx, log_n, TW_COSET_FWD, st);
  const FrEl* tw_ci = ntt_twiddles(ctx, log_n, TW_COSET_INV, st);
  for (int i = 1; i < p.npass; i++) {
    PassArgs pa = pass_args(D.x[0], D.x[1], D.x[2], tw_cf, nullptr, p.pass[i]);
    dist_fill(pa, D, obC, kNoDist, nullptr, nullptr, nullptr);
    launch_pass<PASS_DIT>(ctx, pa, ll, 3, st);
  }
  for (int i = p.npass - 1; i >= 1; i--) {
    PassArgs pa = pass_args(D.x[0], D.x[1], D.x[2], tw_ci, nullptr, p.pass[i]);
    dist_fill(pa, D, obC, i == 1 ? obR : kNoDist, D.peer_y[0], nullptr, nullptr);
    if (i == p.npass - 1) { pa.flags |= PF_POINTWISE_IN; pa.k1 = d.wm_k1; pa.k2 = d.wm_k2; }
    launch_pass<PASS_DIF>(ctx, pa, ll, 1, st);
  }
}
// step 4: coset-inverse low pass on Y_a: this rank's chunk [n me / W, n (me + 1) / W) of h in bit-reversed
// order, canonical
void wm_dist_step4(Ctx* ctx, const NttDist& D, uint32_t log_n, cudaStream_t st) {
  const Plan p = make_plan(log_n, D.wbits);
  const uint32_t obR = log_n - D.wbits, ll = log_n - D.wbits;
  PassArgs pa = pass_args(D.y[0], nullptr, nullptr, ntt_twiddles(ctx, log_n, TW_COSET_INV, st), nullptr, p.pass[0]);
  dist_fill(pa, D, obR, kNoDist, nullptr, nullptr, nullptr);
  pa.flags |= PF_CANONICAL_OUT;
  launch_pass<PASS_DIF>(ctx, pa, ll, 1, st);
}

// ---------------------------------------------------------------------------
// Profiling spans
// ---------------------------------------------------------------------------
ProfileScope::ProfileScope(Ctx* c, int phase, cudaStream_t s, uint64_t units, const uint32_t* d_units)
    : ctx(c), st(s) {
  if (!c->profile) return;
  ProfileSpan sp;
  sp.phase = phase;
  sp.units = units;
  sp.units_pinned = nullptr;
  auto take_event = [&](cudaEvent_t* e) {
    if (!c->event_pool.empty()) { *e = c->event_pool.back(); c->event_pool.pop_back(); }
    else B2Z_CUDA(cudaEventCreate(e));
  };
  take_event(&sp.start);
  take_event(&sp.stop);
  if (d_units != nullptr && c->units_pool != nullptr && c->units_used < kUnitsPool) {
    sp.units_pinned = c->units_pool + c->units_used++;
    *sp.units_pinned = 0;
    // the count is produced by an earlier kernel on the same stream
    B2Z_CUDA(cudaMemcpyAsync(sp.units_pinned, d_units, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  }
  B2Z_CUDA(cudaEventRecord(sp.start, s));
  std::lock_guard<std::mutex> lock(c->span_mu);
  c->spans.push_back(sp);
  idx = (int)c->spans.size() - 1;
}

ProfileScope::~ProfileScope() {
  if (idx < 0) return;
  cudaEventRecord(ctx->spans[idx].stop, st);
}

// ---------------------------------------------------------------------------
// Integer-pipe peak: register-resident multiply-add chains at full occupancy.
// Gives the denominators of the MSM / NTT integer rooflines.
// ---------------------------------------------------------------------------
namespace {
constexpr int kPeakIters = 4096, kPeakChains = 8;

__global__ void __launch_bounds__(256) imad_peak_kernel(uint32_t* out, uint32_t a, uint32_t b) {
  uint32_t x[kPeakChains];
#pragma unroll
  for (int j = 0; j < kPeakChains; j++) x[j] = threadIdx.x + j;
#pragma unroll 1
  for (int i = 0; i < kPeakIters; i++) {
#pragma unroll
    for (int j = 0; j < kPeakChains; j++) x[j] = x[j] * a + b;     // IMAD
  }
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < kPeakChains; j++) s ^= x[j];
  if (s == 0x12345678u) out[0] = s;
}

// the instruction the field multiplications are made of: (lo, hi) += a*b with carry in/out
__global__ void __launch_bounds__(256) imad_wide_peak_kernel(uint32_t* out, uint32_t a, uint32_t b) {
  uint32_t lo[kPeakChains], hi[kPeakChains];
#pragma unroll
  for (int j = 0; j < kPeakChains; j++) { lo[j] = threadIdx.x + j; hi[j] = j; }
  ptx::CF cf;
#pragma unroll 1
  for (int i = 0; i < kPeakIters; i++) {
    ptx::mad_wide_cc(cf, lo[0], hi[0], a, b);
#pragma unroll
    for (int j = 1; j < kPeakChains; j++) ptx::madc_wide_cc(cf, lo[j], hi[j], a, b);
  }
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < kPeakChains; j++) s ^= lo[j] ^ hi[j];
  if (s == 0x12345678u) out[0] = s;
}
}  // namespace

void measure_int_peak(Ctx* ctx, double* imad_per_s, double* imad_wide_per_s) {
  cudaStream_t st = ctx->stream;
  int sms = 0;
  B2Z_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
  DevBuf<uint32_t> out(1);
  const int grid = sms * 8, block = 256;
  cudaEvent_t e0, e1;
  B2Z_CUDA(cudaEventCreate(&e0));
  B2Z_CUDA(cudaEventCreate(&e1));
  double best[2] = {0, 0};
  for (int which = 0; which < 2; which++) {
    for (int rep = 0; rep < 5; rep++) {
      B2Z_CUDA(cudaEventRecord(e0, st));
      if (which == 0) imad_peak_kernel<<<grid, block, 0, st>>>(out.p, 0x9e3779b1u + rep, 0x7f4a7c15u);
      else imad_wide_peak_kernel<<<grid, block, 0, st>>>(out.p, 0x9e3779b1u + rep, 0x7f4a7c15u);
      B2Z_LAUNCHED(ctx);
      B2Z_CUDA(cudaEventRecord(e1, st));
      B2Z_CUDA(cudaEventSynchronize(e1));
      float ms = 0;
      B2Z_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      const double ops = (double)grid * block * kPeakIters * kPeakChains;
      const double rate = ops / (ms * 1e-3);
      if (rep > 0 && rate > best[which]) best[which] = rate;
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *imad_per_s = best[0];
  *imad_wide_per_s = best[1];
}

}  // namespace b2z
