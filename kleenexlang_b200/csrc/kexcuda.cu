// kexcuda.cu -- B200 (sm_100a) execution back end for compiled Kleenex
// streaming string transducers.  C ABI in include/kexcuda.h.
//
// What the reference does sequentially (one goto-threaded `matchN()` per
// phase, src/KMC/Program/Backends/C.hs:72-83, over the buffered runtime
// crt/crt.c) is evaluated here as a prefix computation over the SST's
// transition monoid (src/KMC/SymbolicSST.hs:122-136):
//
// Three kernel families implement the same pipeline behind the C ABI; kex_load
// picks the fastest one whose packed table fields fit the program:
//   kex_v3.cuh    k3_fwd / k3_seams / k3_emit: monoid tables, warp-autonomous
//                 emit (per-lane replicated tables, scan warp, swizzled staging)
//   kex_fast.cuh  k_fwd_monoid / k_seams / k_emit_fast: monoid tables, CTA tiles
//   this file     the generic kernels below (any copyless SST <= 32 registers):
//
//   k_chunk_maps   K1  per 4 KiB chunk: state map Q -> Q, speculating over all
//                      start states with converging tracks   (readnext/consume
//                      + the state chain of matchN, run for every start state)
//   k_compose /    K2  hierarchical composition of the maps and push-down of
//   k_push_states      the true start state of every chunk  (the sequential
//                      dependence between crt.c's 16 KiB windows)
//   k_true_walk    K3  per chunk from its true start state: state samples
//                      every 32 B, first failing position (C.hs:79-81),
//                      register fate map / resolved + pending byte counts
//                      (crt.c:161-259 append/concat/reset, counted not copied)
//   k_compose /    K4  backward composition of the fate maps: which registers
//   k_push_live        at a chunk's end still reach the stream; output offsets
//   k_emit         K5  per chunk: decide per created byte whether it survives,
//                      then write it at its final offset (outputconst /
//                      outputarray / output, crt.c:217-283) through a shared-
//                      memory staging window and 16-byte coalesced stores
//
// The chunk that creates a byte writes it; registers never move data.
// Integer/byte work only: no tensor cores.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/kexcuda.h"

#define KEX_CHUNK 4096u          // bytes per chunk (K1/K3 thread, K5 CTA)
#define KEX_SUB 32u              // bytes per K5 thread
#define KEX_NT (KEX_CHUNK / KEX_SUB)   // K5 threads per CTA (128)
#define KEX_FANIN 64u            // maps composed per group per level
#define KEX_TRACKS 8             // speculative tracks kept in registers (K1)
#define KEX_STAGE 20480u         // K5 staging window bytes
#define KEX_DEAD 0xFFu
#define KEX_NONE32 0xFFFFFFFFu
#define KEX_NONE64 0xFFFFFFFFFFFFFFFFull

enum { KIND_NOP = 0, KIND_OUTSYM = 1, KIND_OUTONLY = 2, KIND_GENERAL = 3 };
enum { PIECE_CONST = 0, PIECE_SYM = 1 };

struct ActHdr { uint32_t kind, npieces, piece_off, outlen0, flush_mask, total_len; };
struct Piece { uint8_t target, kind; uint16_t len; uint32_t off; };

struct PhaseDev {
  uint32_t Q, C, R, A, init, max_out;
  const uint8_t *cls;        // [256]
  const uint32_t *trans;     // [(Q+1)*C]  next | action << 16
  const uint16_t *nextpm;    // [(Q+1)*C]  next state pre-multiplied by C
  const ActHdr *acts;        // [A]
  const uint32_t *actinfo;   // [A] kind | outlen0 << 8
  const uint8_t *fate;       // [A*R]
  const uint32_t *addlen;    // [A*R]
  const Piece *pieces;
  const uint8_t *consts;
  const uint32_t *img_cnt;   // [C]   distinct non-FAIL successors per class
  const uint16_t *img_state; // [C*Q] those successors (pre-multiplied by C)
  const uint16_t *img_idx;   // [C*(Q+1)] start state -> track index or 0xFFFF
};

struct RunResult {               // device -> host summary of a walk
  unsigned long long fail_pos;
  unsigned long long total_out;
  uint32_t end_state;
  uint32_t pad;
};

// ---------------------------------------------------------------- helpers
__device__ __forceinline__ uint4 ld_stream16(const uint8_t *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ uint32_t byte_of(const uint4 &v, int j) {
  uint32_t w = (j < 4) ? v.x : (j < 8) ? v.y : (j < 12) ? v.z : v.w;
  return (w >> ((j & 3) * 8)) & 0xFFu;
}

// ------------------------------------------------------------------- K1
// One thread per chunk.  Track t follows the t-th distinct successor of the
// chunk's first byte; tracks that meet stay together, so the loop collapses
// to a single dependent lookup per byte once they have converged.
__global__ void __launch_bounds__(128)
k_chunk_maps(PhaseDev P, const uint8_t *__restrict__ in, size_t n, size_t nchunks,
             uint16_t *__restrict__ maps) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint8_t *cls = smem;
  uint16_t *nxt = (uint16_t *)(smem + 256);
  const uint32_t Q1 = P.Q + 1, C = P.C;
  for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) cls[i] = P.cls[i];
  for (uint32_t i = threadIdx.x; i < Q1 * C; i += blockDim.x) nxt[i] = P.nextpm[i];
  __syncthreads();
  size_t chunk = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (chunk >= nchunks) return;
  const size_t base = chunk * KEX_CHUNK;
  const uint32_t len = (uint32_t)((n - base < KEX_CHUNK) ? (n - base) : KEX_CHUNK);
  const uint32_t FAILPM = P.Q * C;
  uint16_t *row = maps + chunk * Q1;
  const uint8_t *p = in + base;
  const uint32_t nwords = (len + 15u) >> 4;
  uint4 w0 = ld_stream16(p);
  const uint32_t c0 = cls[w0.x & 0xFFu];
  const uint32_t m = P.img_cnt[c0];
  const uint16_t *ist = P.img_state + (size_t)c0 * P.Q;
  const uint16_t *iidx = P.img_idx + (size_t)c0 * Q1;
  for (uint32_t q = 0; q < Q1; ++q) row[q] = (uint16_t)P.Q;
  for (uint32_t tb = 0; tb < m; tb += KEX_TRACKS) {
    uint32_t s[KEX_TRACKS];
#pragma unroll
    for (int t = 0; t < KEX_TRACKS; ++t) s[t] = (tb + t < m) ? ist[tb + t] : FAILPM;
    bool single = false;
    uint32_t alive = 0, s0 = FAILPM;
    for (uint32_t w = 0; w < nwords; ++w) {
      uint4 v = (w == 0) ? w0 : ld_stream16(p + 16u * w);
      const uint32_t lim = (len - 16u * w < 16u) ? (len - 16u * w) : 16u;
      if (single) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if ((w == 0 && j == 0) || (uint32_t)j >= lim) continue;
          s0 = nxt[s0 + cls[byte_of(v, j)]];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if ((w == 0 && j == 0) || (uint32_t)j >= lim) continue;
          const uint32_t c = cls[byte_of(v, j)];
#pragma unroll
          for (int t = 0; t < KEX_TRACKS; ++t) s[t] = nxt[s[t] + c];
        }
        // converged?  all live tracks in one state
        uint32_t v0 = FAILPM;
        bool conv = true;
#pragma unroll
        for (int t = 0; t < KEX_TRACKS; ++t) {
          if (s[t] != FAILPM) {
            if (v0 == FAILPM) v0 = s[t];
            else if (s[t] != v0) conv = false;
          }
        }
        if (conv) {
          single = true;
          s0 = v0;
          alive = 0;
#pragma unroll
          for (int t = 0; t < KEX_TRACKS; ++t) if (s[t] != FAILPM) alive |= 1u << t;
          if (v0 == FAILPM) break;   // every track of this batch is dead
        }
      }
    }
    if (single) {
#pragma unroll
      for (int t = 0; t < KEX_TRACKS; ++t) s[t] = ((alive >> t) & 1u) ? s0 : FAILPM;
    }
    for (uint32_t q = 0; q < P.Q; ++q) {
      const uint32_t ix = iidx[q];
      if (ix >= tb && ix < tb + KEX_TRACKS) {
        uint32_t r = FAILPM;
#pragma unroll
        for (int t = 0; t < KEX_TRACKS; ++t) if (ix - tb == (uint32_t)t) r = s[t];
        row[q] = (uint16_t)(r / C);
      }
    }
  }
}

// ------------------------------------------------------------- K2 / K4 scans
// parent[g] = child[g*G] ; child[g*G+1] ; ...   (apply left to right).
// FATE maps have absorbing values 0 (flushed) and 0xFF (dropped).
template <typename T, bool FATE>
__global__ void k_compose(const T *__restrict__ child, size_t nchild, T *__restrict__ parent,
                          uint32_t D) {
  const size_t g = blockIdx.x;
  const size_t lo = g * KEX_FANIN;
  const size_t hi = (lo + KEX_FANIN < nchild) ? (lo + KEX_FANIN) : nchild;
  for (uint32_t q = threadIdx.x; q < D; q += blockDim.x) {
    uint32_t s = q;
    for (size_t j = lo; j < hi; ++j) {
      if (FATE && (s == 0u || s == KEX_DEAD)) break;
      s = child[j * D + s];
    }
    parent[g * D + q] = (T)s;
  }
}

__global__ void k_push_states(const uint16_t *__restrict__ child_maps, size_t nchild,
                              const uint16_t *__restrict__ parent_start, size_t nparent,
                              uint16_t *__restrict__ child_start, uint32_t D) {
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nparent) return;
  const size_t lo = g * KEX_FANIN;
  const size_t hi = (lo + KEX_FANIN < nchild) ? (lo + KEX_FANIN) : nchild;
  uint32_t s = parent_start[g];
  for (size_t j = lo; j < hi; ++j) {
    child_start[j] = (uint16_t)s;
    s = child_maps[j * D + s];
  }
}

__device__ __forceinline__ uint32_t pullback(const uint8_t *__restrict__ F, uint32_t R, uint32_t L) {
  // registers whose content goes (through F) to the stream or to a live register
  const uint32_t m = L | 1u;
  uint32_t r = 0;
  for (uint32_t i = 1; i < R; ++i) {
    const uint32_t d = F[i];
    if (d != KEX_DEAD && ((m >> d) & 1u)) r |= 1u << i;
  }
  return r;
}

__global__ void k_push_live(const uint8_t *__restrict__ child_fate, size_t nchild,
                            const uint32_t *__restrict__ parent_live, size_t nparent,
                            uint32_t *__restrict__ child_live, uint32_t R) {
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nparent) return;
  const size_t lo = g * KEX_FANIN;
  const size_t hi = (lo + KEX_FANIN < nchild) ? (lo + KEX_FANIN) : nchild;
  uint32_t L = parent_live[g];
  for (size_t j = hi; j-- > lo;) {
    child_live[j] = L;
    L = pullback(child_fate + j * R, R, L);
  }
}

__global__ void k_set_u16(uint16_t *p, uint32_t v) { *p = (uint16_t)v; }
__global__ void k_set_u32(uint32_t *p, uint32_t v) { *p = v; }

// ------------------------------------------------------------------- K3
__global__ void __launch_bounds__(128)
k_true_walk(PhaseDev P, const uint8_t *__restrict__ in, size_t n, size_t nchunks,
            const uint16_t *__restrict__ start, uint16_t *__restrict__ samples,
            uint8_t *__restrict__ chunk_fate, uint32_t *__restrict__ chunk_pend,
            unsigned long long *__restrict__ chunk_resolved, uint32_t *__restrict__ chunk_fail,
            RunResult *__restrict__ res) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint8_t *cls = smem;
  const uint32_t Q1 = P.Q + 1, C = P.C, R = P.R;
  uint32_t *trans = (uint32_t *)(smem + 256);
  uint32_t *actinfo = trans + Q1 * C;
  for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) cls[i] = P.cls[i];
  for (uint32_t i = threadIdx.x; i < Q1 * C; i += blockDim.x) trans[i] = P.trans[i];
  for (uint32_t i = threadIdx.x; i < P.A; i += blockDim.x) actinfo[i] = P.actinfo[i];
  __syncthreads();
  size_t chunk = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (chunk >= nchunks) return;
  const size_t base = chunk * KEX_CHUNK;
  const uint32_t len = (uint32_t)((n - base < KEX_CHUNK) ? (n - base) : KEX_CHUNK);
  const uint8_t *p = in + base;
  uint32_t s = start[chunk];
  uint8_t F[32];
  uint32_t ln[32];
  for (uint32_t r = 0; r < R; ++r) { F[r] = (uint8_t)r; ln[r] = 0; }
  unsigned long long resolved = 0;
  uint32_t fail = KEX_NONE32;
  uint16_t *srow = samples + chunk * KEX_NT;
  const uint32_t nwords = (len + 15u) >> 4;
  uint32_t pk[4] = {0, 0, 0, 0};
  uint32_t nrec = 0;   // samples recorded so far
  for (uint32_t w = 0; w < nwords && fail == KEX_NONE32; ++w) {
    uint4 v = ld_stream16(p + 16u * w);
    const uint32_t lim = (len - 16u * w < 16u) ? (len - 16u * w) : 16u;
    if ((w & 1u) == 0) {   // a 32-byte sub-chunk starts here
      const uint32_t slot = nrec & 7u;
      if (slot & 1u) pk[slot >> 1] |= s << 16; else pk[slot >> 1] = s;
      ++nrec;
      if (slot == 7u) *(uint4 *)(srow + (nrec - 8u)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if ((uint32_t)j >= lim || fail != KEX_NONE32) continue;
      const uint32_t e = trans[s * C + cls[byte_of(v, j)]];
      const uint32_t ns = e & 0xFFFFu, a = e >> 16;
      if (ns == P.Q) { fail = 16u * w + j; continue; }
      s = ns;
      const uint32_t info = actinfo[a];
      if ((info & 0xFFu) != KIND_GENERAL) {
        resolved += info >> 8;
      } else {
        const uint8_t *fa = P.fate + (size_t)a * R;
        const uint32_t *al = P.addlen + (size_t)a * R;
        uint32_t nl[32];
        for (uint32_t r = 0; r < R; ++r) nl[r] = al[r];
        for (uint32_t r = 1; r < R; ++r) {
          const uint32_t d = fa[r];
          if (d == 0u) resolved += ln[r];
          else if (d != KEX_DEAD) nl[d] += ln[r];
        }
        resolved += nl[0];
        for (uint32_t r = 1; r < R; ++r) ln[r] = nl[r];
        for (uint32_t r = 1; r < R; ++r) {
          const uint32_t x = F[r];
          if (x != 0u && x != KEX_DEAD) F[r] = fa[x];
        }
      }
    }
  }
  for (uint32_t k = nrec & ~7u; k < nrec; ++k) {   // partially filled sample pack
    const uint32_t slot = k & 7u;
    srow[k] = (uint16_t)((slot & 1u) ? (pk[slot >> 1] >> 16) : (pk[slot >> 1] & 0xFFFFu));
  }
  chunk_fail[chunk] = fail;
  chunk_resolved[chunk] = resolved;
  if (R > 1) {
    for (uint32_t r = 0; r < R; ++r) {
      chunk_fate[chunk * R + r] = F[r];
      chunk_pend[chunk * R + r] = ln[r];
    }
  }
  if (chunk == nchunks - 1) res->end_state = (fail == KEX_NONE32) ? s : P.Q;
}

__global__ void k_reduce_fail(const uint32_t *__restrict__ chunk_fail, size_t nchunks,
                              RunResult *__restrict__ res) {
  __shared__ unsigned long long best[32];
  unsigned long long m = KEX_NONE64;
  for (size_t i = threadIdx.x; i < nchunks; i += blockDim.x) {
    const uint32_t f = chunk_fail[i];
    if (f != KEX_NONE32) {
      const unsigned long long pos = (unsigned long long)i * KEX_CHUNK + f;
      if (pos < m) m = pos;
      break;   // later chunks of this thread are further right
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long x = __shfl_xor_sync(0xFFFFFFFFu, m, o);
    if (x < m) m = x;
  }
  if ((threadIdx.x & 31) == 0) best[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (uint32_t w = 1; w < (blockDim.x >> 5); ++w) if (best[w] < m) m = best[w];
    res->fail_pos = m;
  }
}

// -------------------------------------------------- output lengths & offsets
__global__ void k_outlen(const uint32_t *__restrict__ pend, const unsigned long long *__restrict__ resolved,
                         const uint32_t *__restrict__ live, size_t nchunks, uint32_t R,
                         unsigned long long *__restrict__ outlen) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nchunks) return;
  unsigned long long t = resolved[i];
  if (R > 1) {
    const uint32_t L = live[i];
    for (uint32_t r = 1; r < R; ++r) if ((L >> r) & 1u) t += pend[i * R + r];
  }
  outlen[i] = t;
}

#define SCAN_ITEMS 8
#define SCAN_THREADS 256
#define SCAN_TILE (SCAN_ITEMS * SCAN_THREADS)

__device__ __forceinline__ unsigned long long block_excl_scan(unsigned long long v, unsigned long long *total,
                                                              unsigned long long *sh) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long x = v;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, x, o);
    if (lane >= (uint32_t)o) x += y;
  }
  if (lane == 31) sh[warp] = x;
  __syncthreads();
  if (warp == 0) {
    unsigned long long w = (lane < (blockDim.x >> 5)) ? sh[lane] : 0ull;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, w, o);
      if (lane >= (uint32_t)o) w += y;
    }
    sh[32 + lane] = w;   // inclusive warp totals
  }
  __syncthreads();
  const unsigned long long wbase = warp ? sh[32 + warp - 1] : 0ull;
  *total = sh[32 + (blockDim.x >> 5) - 1];
  const unsigned long long r = wbase + x - v;
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_sums(const unsigned long long *__restrict__ v, size_t n, unsigned long long *__restrict__ bsum) {
  __shared__ unsigned long long sh[64];
  const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  unsigned long long t = 0;
  for (int k = 0; k < SCAN_ITEMS; ++k) if (base + k < n) t += v[base + k];
  unsigned long long total;
  block_excl_scan(t, &total, sh);
  if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_top(unsigned long long *__restrict__ bsum, size_t nb, RunResult *__restrict__ res) {
  __shared__ unsigned long long sh[64];
  unsigned long long carry = 0;
  for (size_t base = 0; base < nb; base += SCAN_THREADS) {
    const size_t i = base + threadIdx.x;
    const unsigned long long v = (i < nb) ? bsum[i] : 0ull;
    unsigned long long total;
    const unsigned long long e = block_excl_scan(v, &total, sh);
    if (i < nb) bsum[i] = carry + e;
    carry += total;
  }
  if (threadIdx.x == 0) res->total_out = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_apply(const unsigned long long *__restrict__ v, size_t n, const unsigned long long *__restrict__ bsum,
             unsigned long long *__restrict__ off) {
  __shared__ unsigned long long sh[64];
  const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  unsigned long long loc[SCAN_ITEMS];
  unsigned long long t = 0;
  for (int k = 0; k < SCAN_ITEMS; ++k) { loc[k] = (base + k < n) ? v[base + k] : 0ull; t += loc[k]; }
  unsigned long long total;
  unsigned long long e = block_excl_scan(t, &total, sh) + bsum[blockIdx.x];
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (base + k < n) off[base + k] = e;
    e += loc[k];
  }
}

// ------------------------------------------------------------------- K5
// One CTA per chunk, one thread per 32-byte sub-chunk.
//   pass A  forward from the sampled state: action id per position, bytes that
//           certainly reach the stream, fate map of the sub-chunk
//   scan    suffix composition of the sub-chunk fate maps -> live registers at
//           each sub-chunk's end (only if some position moves registers)
//   pass B  backward: survival mask per register-moving position, byte count
//   pass C  forward: write surviving bytes at their final offsets
// MB = bytes of a survival mask kept per position (0: program has no registers)
template <int MB>
__global__ void __launch_bounds__(KEX_NT)
k_emit(PhaseDev P, const uint8_t *__restrict__ in, size_t n_eff,
       const uint16_t *__restrict__ samples, const unsigned long long *__restrict__ outoff,
       const uint32_t *__restrict__ live_end, uint8_t *__restrict__ out) {
  extern __shared__ __align__(16) uint8_t smem[];
  const uint32_t Q1 = P.Q + 1, C = P.C, R = P.R;
  const uint32_t tid = threadIdx.x;
  // ---- carve shared memory
  uint8_t *in_sm = smem;                                   // KEX_CHUNK
  uint8_t *stage = in_sm + KEX_CHUNK;                      // KEX_STAGE + 32
  uint16_t *act_sm = (uint16_t *)(stage + KEX_STAGE + 32); // KEX_CHUNK entries, [j][tid]
  uint8_t *mask_sm = (uint8_t *)(act_sm + KEX_CHUNK);      // KEX_CHUNK * MB
  uint8_t *maps_sm = mask_sm + (size_t)KEX_CHUNK * MB;     // 2 * KEX_NT * R   (MB > 0)
  uint8_t *tbl = maps_sm + (MB ? 2u * KEX_NT * 32u : 0u);
  uint32_t *trans = (uint32_t *)tbl;
  uint32_t *actinfo = trans + Q1 * C;
  uint8_t *cls = (uint8_t *)(actinfo + P.A);
  __shared__ uint32_t warp_sums[KEX_NT / 32];
  __shared__ uint32_t tile_total;

  for (uint32_t i = tid; i < Q1 * C; i += KEX_NT) trans[i] = P.trans[i];
  for (uint32_t i = tid; i < P.A; i += KEX_NT) actinfo[i] = P.actinfo[i];
  for (uint32_t i = tid; i < 256; i += KEX_NT) cls[i] = P.cls[i];

  const size_t tile = blockIdx.x;
  const size_t base = tile * KEX_CHUNK;
  const uint32_t tlen = (uint32_t)((n_eff - base < KEX_CHUNK) ? (n_eff - base) : KEX_CHUNK);
  {
    const uint32_t nw = (tlen + 15u) >> 4;
    for (uint32_t w = tid; w < nw; w += KEX_NT)
      *(uint4 *)(in_sm + 16u * w) = ld_stream16(in + base + 16u * w);
  }
  __syncthreads();

  // ---- pass A
  const uint32_t lo = tid * KEX_SUB;
  const uint32_t cnt_pos = (lo < tlen) ? ((tlen - lo < KEX_SUB) ? (tlen - lo) : KEX_SUB) : 0u;
  uint32_t s = cnt_pos ? samples[tile * KEX_NT + tid] : 0u;
  uint32_t cnt = 0;
  bool anygen = false;
  uint8_t f[32];
  if (MB) for (uint32_t r = 0; r < R; ++r) f[r] = (uint8_t)r;
  for (uint32_t j = 0; j < cnt_pos; ++j) {
    const uint32_t e = trans[s * C + cls[in_sm[lo + j]]];
    const uint32_t a = e >> 16;
    s = e & 0xFFFFu;
    act_sm[j * KEX_NT + tid] = (uint16_t)a;
    const uint32_t info = actinfo[a];
    if (!MB || (info & 0xFFu) != KIND_GENERAL) {
      cnt += info >> 8;
    } else {
      anygen = true;
      const uint8_t *fa = P.fate + (size_t)a * R;
      for (uint32_t r = 1; r < R; ++r) {
        const uint32_t x = f[r];
        if (x != 0u && x != KEX_DEAD) f[r] = fa[x];
      }
    }
  }

  if (MB) {
    const int any = __syncthreads_or(anygen ? 1 : 0);
    if (any) {
      // ---- suffix composition of the sub-chunk fate maps (Hillis-Steele)
      uint8_t *cur = maps_sm, *nxt = maps_sm + KEX_NT * R;
      for (uint32_t r = 0; r < R; ++r) cur[tid * R + r] = f[r];
      __syncthreads();
      for (uint32_t d = 1; d < KEX_NT; d <<= 1) {
        for (uint32_t r = 0; r < R; ++r) {
          uint32_t x = cur[tid * R + r];
          if (tid + d < KEX_NT && x != 0u && x != KEX_DEAD) x = cur[(tid + d) * R + x];
          nxt[tid * R + r] = (uint8_t)x;
        }
        __syncthreads();
        uint8_t *t = cur; cur = nxt; nxt = t;
      }
      // ---- pass B
      if (anygen) {
        const uint32_t Lt = live_end[tile];
        uint32_t L = (tid == KEX_NT - 1) ? Lt : pullback(cur + (tid + 1) * R, R, Lt);
        for (uint32_t j = cnt_pos; j-- > 0;) {
          const uint32_t a = act_sm[j * KEX_NT + tid];
          if ((actinfo[a] & 0xFFu) != KIND_GENERAL) continue;
          const uint32_t m = L | 1u;
          if (MB == 1) mask_sm[j * KEX_NT + tid] = (uint8_t)m;
          else ((uint32_t *)mask_sm)[j * KEX_NT + tid] = m;
          const uint8_t *fa = P.fate + (size_t)a * R;
          const uint32_t *al = P.addlen + (size_t)a * R;
          uint32_t nl = 0;
          for (uint32_t r = 0; r < R; ++r) {
            if ((m >> r) & 1u) cnt += al[r];
            const uint32_t d = fa[r];
            if (r && d != KEX_DEAD && ((m >> d) & 1u)) nl |= 1u << r;
          }
          L = nl;
        }
      }
    }
  }

  // ---- exclusive scan of per-thread byte counts
  uint32_t x = cnt;
  const uint32_t lane = tid & 31, warp = tid >> 5;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
    if (lane >= (uint32_t)o) x += y;
  }
  if (lane == 31) warp_sums[warp] = x;
  __syncthreads();
  if (tid == 0) {
    uint32_t t = 0;
    for (uint32_t w = 0; w < KEX_NT / 32; ++w) { const uint32_t v = warp_sums[w]; warp_sums[w] = t; t += v; }
    tile_total = t;
  }
  __syncthreads();
  uint32_t o = warp_sums[warp] + x - cnt;
  const uint32_t total = tile_total;
  const unsigned long long gbase = outoff[tile];
  const uint32_t shift = (uint32_t)(gbase & 15ull);
  const bool staged = (total + shift) <= KEX_STAGE;

  // ---- pass C
  for (uint32_t j = 0; j < cnt_pos; ++j) {
    const uint32_t a = act_sm[j * KEX_NT + tid];
    const uint32_t kind = actinfo[a] & 0xFFu;
    if (kind == KIND_NOP) continue;
    const uint8_t sym = in_sm[lo + j];
    if (kind == KIND_OUTSYM) {
      if (staged) stage[shift + o] = sym; else out[gbase + o] = sym;
      ++o;
      continue;
    }
    uint32_t m = 1u;
    if (MB && kind == KIND_GENERAL)
      m = (MB == 1) ? (uint32_t)mask_sm[j * KEX_NT + tid] : ((const uint32_t *)mask_sm)[j * KEX_NT + tid];
    const ActHdr h = P.acts[a];
    for (uint32_t k = 0; k < h.npieces; ++k) {
      const Piece pc = P.pieces[h.piece_off + k];
      if (!((m >> pc.target) & 1u)) continue;
      if (pc.kind == PIECE_SYM) {
        if (staged) stage[shift + o] = sym; else out[gbase + o] = sym;
        ++o;
      } else {
        const uint8_t *c = P.consts + pc.off;
        for (uint32_t t = 0; t < pc.len; ++t) {
          if (staged) stage[shift + o + t] = c[t]; else out[gbase + o + t] = c[t];
        }
        o += pc.len;
      }
    }
  }
  if (!staged) return;
  __syncthreads();
  // ---- staging window -> global, 16-byte words aligned to the destination
  uint8_t *gal = out + (gbase - shift);
  const uint32_t end = shift + total;
  const uint32_t w_lo = shift ? 1u : 0u;
  const uint32_t w_hi = end >> 4;
  for (uint32_t w = w_lo + tid; w < w_hi; w += KEX_NT)
    *(uint4 *)(gal + 16u * w) = *(const uint4 *)(stage + 16u * w);
  if (shift) {
    const uint32_t he = (end < 16u) ? end : 16u;
    for (uint32_t b = shift + tid; b < he; b += KEX_NT) gal[b] = stage[b];
  }
  if (w_hi >= w_lo) {
    for (uint32_t b = (w_hi << 4) + tid; b < end; b += KEX_NT)
      if (b >= shift) gal[b] = stage[b];
  }
}

__global__ void k_copy_tail(const uint8_t *__restrict__ src, uint32_t len, uint8_t *__restrict__ dst) {
  for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) dst[i] = src[i];
}

// small results (run summary, scan control block, state map, seam summary) go to
// the host through mapped pinned memory instead of a device->host copy: a copy
// would queue behind the output copies of kex_run_host's pipeline on the copy engine
__global__ void k_publish(const uint8_t *__restrict__ src, volatile uint8_t *dst, uint32_t n) {
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
  __threadfence_system();
}

#include "kex_fast.cuh"
#include "kex_v3.cuh"
#include "kex_v4.cuh"
#include "kex_act.cuh"

// =================================================================== host
struct PhaseHost {
  PhaseDev dev;
  std::vector<int32_t> fin;          // final action per state or -1
  std::vector<ActHdr> acts;
  std::vector<uint8_t> fate;
  std::vector<Piece> pieces;
  std::vector<uint8_t> consts;
  void *d_blob = nullptr;            // device copy of the phase blob
  void *d_extra = nullptr;           // derived tables (nextpm, actinfo, img_*)
  size_t smem_walk = 0, smem_maps = 0, smem_emit = 0;
  int mask_bytes = 0;
  // monoid kernels (fast section of the blob); absent -> generic kernels
  bool fast = false;
  FastDev fdev;
  std::vector<uint8_t> lam_final;     // [Q+1] live-set index flushed at end of input, 0xFF = reject
  std::vector<uint32_t> lam_masks;    // [NL] register mask of every live set
  size_t smem_fm = 0, smem_ef_tables = 0;
  uint32_t stage_bytes = 12288;       // emit staging window; grows with the observed out/in ratio
  int ef_ctas_per_sm = 0;
  uint32_t ef_stage_cfg = 0;          // stage_bytes the occupancy was computed for
  // warp-autonomous kernels (kex_v3.cuh); absent -> kex_fast.cuh kernels
  V3Dev v3;
  void *d_v3 = nullptr;               // be3, tpl2, fwdtab
  size_t smem_fwd3 = 0, smem_seams3 = 0;
  uint32_t v3_stage = 3072;           // per-warp staging window; follows the observed out/in ratio
  uint32_t v3_reccap = V3_RECCAP;     // template records per tile kept in shared memory; follows the observed maximum
  // G-mode emit kernel (kex_v4.cuh); absent / switched off -> k3_emit
  V4Dev v4;
  void *d_v4 = nullptr;               // gtab, G, tpl, apply8
  std::vector<uint32_t> h_trans2, h_BE, g_static_tab;   // host copies for learn_gmode
  std::vector<uint8_t> g_static;      // [Q+1] the blob's static choice of G
  bool v4_learned = false, v4_off = false;
  bool v4_tail_ok = true;             // tail evaluation has not had to be repeated exactly so far
  std::vector<uint8_t> h_G;           // host copy of the G the device tables were built for
  uint32_t v4_relearns = 0, v4_last_exact = 0;
  uint32_t v4_stage = 3072, v4_reccap = V4_RECCAP;
  // action-interpreter phase (kex_act.cuh): no SST tables at all
  bool act = false;
  uint32_t act_nregs = 0;
  size_t tile() const { return v3.ok ? (size_t)V3_TILE : (size_t)KEX_CHUNK; }
};

struct Buf {
  void *p = nullptr;
  size_t cap = 0;
};

// scratch of the action-interpreter kernels
struct ActScratch {
  Buf delta, mn, mx, h0, bsum, fate, add, gfate, gadd, gvec, vec, wlen, ctl;
};

// Scratch of one run / shard in flight (grow-only device buffers and the state
// carried between the shard steps).  A program holds two so that the host
// pipeline of kex_run_host can walk one sub-wave while the previous one emits.
struct Ctx {
  Buf maps[8], starts[8], fates[8], lives[8];
  Buf samples, blockpre, pend, resolved, fail, outlen, outoff, bsum, res_dev;
  Buf bmaps[8], lams[8], desc, ctl;
  FastCtl *ctl_host = nullptr;
  RunResult *res_host = nullptr;
  uint8_t *hpub = nullptr, *dpub = nullptr;   // 4 KiB of mapped pinned memory (k_publish)
  // shard state between the three shard calls
  const uint8_t *sh_in = nullptr;
  size_t sh_n = 0, sh_nchunks = 0;
  size_t lvl_count[8];
  int nlevels = 0;
  uint32_t sh_phase = 0;
  // G-mode tail evaluation (run_phase): exact live sets exist only for tiles >= tail_first; 0 = for all
  size_t tail_first = 0;
};

struct kex_program {
  int device = 0;
  std::vector<PhaseHost> phases;
  std::string cuda_err;
  uint32_t launches = 0;
  bool timing = false;
  float ms[4] = {0, 0, 0, 0};
  cudaEvent_t ev[8];
  bool ev_ok = false;
  Ctx cx[2];
  Ctx *c = &cx[0];
  Buf inter[2], hostio_in, hostio_out;
  ActScratch as;
  // block streaming (kex_stream_*): two blocks resident, the older one waits for its seam code
  Buf sin[2], sout;
  bool st_open = false, st_failed = false;
  uint32_t st_state = 0, st_blocks = 0, st_nq = 0;      // st_nq: blocks walked but not yet emitted (<= 2)
  size_t st_q[2] = {0, 0}, st_consumed = 0, st_fail_at = 0;
  std::vector<uint8_t> st_seam;                         // seam summary of the newest block when two wait
  int num_sms = 0;
  // host pipeline (kex_run_host)
  cudaStream_t s_h2d = nullptr, s_comp = nullptr, s_d2h = nullptr;
  std::vector<cudaEvent_t> pipe_ev;
  bool shard_tail = false;            // kex_set_shard_tail: kex_shard_walk / kex_shard_emit may use the G-mode tail evaluation
  size_t emit_out_off = 0;            // v3 emit: offset of this shard's output inside d_out
  uint32_t only_phase = 0;            // 1-based phase selected by kex_select_phase, 0 = all
};

#define KEX_NEED_EXACT 100     // internal: repeat the phase with exact live sets for every tile

#define CK(call)                                                              \
  do {                                                                        \
    cudaError_t e_ = (call);                                                  \
    if (e_ != cudaSuccess) {                                                  \
      p->cuda_err = std::string(#call) + ": " + cudaGetErrorString(e_);      \
      return KEX_ERR_CUDA;                                                    \
    }                                                                         \
  } while (0)

static int ensure(kex_program *p, Buf &b, size_t bytes) {
  if (bytes <= b.cap) return KEX_OK;
  if (b.p) CK(cudaFree(b.p));
  b.p = nullptr;
  b.cap = 0;
  size_t want = bytes + bytes / 8 + 256;
  CK(cudaMalloc(&b.p, want));
  b.cap = want;
  return KEX_OK;
}

// device -> host of a few bytes, synchronising the stream
static int fetch_sync(kex_program *p, void *h_dst, const void *d_src, size_t n, cudaStream_t st) {
  if (p->c->hpub && n <= 4096) {
    k_publish<<<1, 128, 0, st>>>((const uint8_t *)d_src, p->c->dpub, (uint32_t)n);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    memcpy(h_dst, p->c->hpub, n);
    return KEX_OK;
  }
  CK(cudaMemcpyAsync(h_dst, d_src, n, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return KEX_OK;
}

static uint32_t rd32(const uint8_t *b, size_t off) {
  uint32_t v;
  memcpy(&v, b + off, 4);
  return v;
}


// Derived tables of the warp-autonomous kernels (kex_v3.cuh).  Programs whose
// tables exceed the kernels' packed fields stay on the kex_fast.cuh kernels.
static int load_v3(kex_program *p, PhaseHost &ph, const uint16_t *mulF, const uint16_t *applyF, const uint32_t *BE,
                   const uint32_t *tplinfo, const uint8_t *poolbytes, uint32_t NM) {
  V3Dev &v = ph.v3;
  memset(&v, 0, sizeof(v));
  if (getenv("KEX_NO_V3")) return KEX_OK;
  const FastDev &F = ph.fdev;
  const uint32_t Q1 = ph.dev.Q + 1, C = ph.dev.C, A = ph.dev.A, NL = F.NL, NB = F.NB, NE = NL * A;
  if ((NL - 1) * A > 255 || NB > 64 || F.NG > 127) return KEX_OK;
  // emission entries
  std::vector<uint32_t> be3(NE), be3w(NE), tpl2(NE, 0);
  bool has_lit = false;
  uint32_t tpl_min = 0xFFFFFFFFu, tpl_max = 0;         // lengths of the templates that get records
  for (uint32_t i = 0; i < NE; ++i) {
    const uint32_t e = BE[i];
    const uint32_t lamA = (e & 0xFFFCu) / 4u, len = (e >> 16) & 0xFFu;      // lam_before * A
    uint32_t S = 0, T = 0, lit = 0;
    if (e & 1u) {
      S = 1;
    } else if (e & 2u) {
      const uint32_t t = e >> 24, src = tplinfo[2 * t] & 0xFFFFu, hm = tplinfo[2 * t + 1];
      if (hm & (hm - 1u)) return KEX_OK;              // more than one hole
      uint32_t hole = 0;
      while (hm >> hole) ++hole;                      // hole offset + 1, 0 = none
      if (len == 1 && hole == 0 && poolbytes[src] >= 1 && poolbytes[src] < 0x80) {
        lit = poolbytes[src];                         // a one-byte ASCII literal is stored by the write pass itself
        has_lit = true;
      } else {
        T = 1;
        tpl2[i] = (src + 4u) | (len << 16) | (hole << 24);      // the shared-memory pool has four bytes of padding in front
        tpl_min = len < tpl_min ? len : tpl_min;
        tpl_max = len > tpl_max ? len : tpl_max;
      }
    }
    be3[i] = len | (T << 15) | (S << 16) | (lamA << 24);
    be3w[i] = be3[i] | (lit << 8);
  }
  if (getenv("KEX_V3_NOLIT") && has_lit) {              // test knob: one-byte literals as ordinary templates
    has_lit = false;
    for (uint32_t i = 0; i < NE; ++i)
      if (be3w[i] & 0x7F00u) {
        const uint32_t e = BE[i], t = e >> 24;
        be3[i] |= 0x8000u;
        tpl2[i] = ((tplinfo[2 * t] & 0xFFFFu) + 4u) | (1u << 16);
        tpl_min = 1;
      }
    be3w = be3;
  }
  // templates of at least three bytes: edge words merged instead of edge bytes (kex_v3.cuh v3_template_rmw);
  // a template of L bytes touches at most (L + 6) / 4 words.  KEX_V3_NORMW: the byte-edge variant (tests)
  v.rmw_words = (tpl_max > 0 && tpl_min >= 3u && !getenv("KEX_V3_NORMW")) ? (tpl_max + 6u) / 4u : 0u;
  // replicated tables must be addressable with 16 bits (2 KiB allowance for the window base)
  uint32_t log = 0;
  for (uint32_t cand : {7u, 6u, 5u}) {                  // one table copy per lane / per two lanes / per four lanes
    if (cand == 6u && getenv("KEX_V3_NOLOG6")) continue;
    const size_t end = (((size_t)NB * 256 + 127) & ~(size_t)127) + ((size_t)Q1 * C + (has_lit ? 2 : 1) * (size_t)NE) * (1u << cand);
    if (end + 2048 <= 65536) { log = cand; break; }
  }
  if (!log) return KEX_OK;
  const uint32_t stride = 1u << log;
  v.log = log;
  v.NE = NE;
  v.has_lit = has_lit ? 1u : 0u;
  uint32_t sp = 0;
  v.o_mulB = sp; sp += NB * 256u;
  sp = (sp + 127u) & ~127u;
  v.o_trans = sp; sp += Q1 * C * stride;
  v.o_BE = sp; sp += (has_lit ? 2u : 1u) * NE * stride;
  sp = (sp + 255u) & ~255u;
  v.o_cls = sp; sp += 256;
  v.o_compB = sp; sp += NB * NB;
  v.o_applyB = sp; sp += NB * NL;
  sp = (sp + 3u) & ~3u;
  v.o_tpl2 = sp; sp += NE * 4u;
  if (F.pool_len + 16u > 0xFFFFu) return KEX_OK;
  v.pool_stride = (F.pool_len + 4u + 7u) & ~3u;
  v.o_pool = sp; sp += 4u * v.pool_stride;
  sp = (sp + 15u) & ~15u;
  v.o_slots = sp; sp += 256u + 512u;
  v.o_guess = sp; sp += (Q1 + 15u) & ~15u;
  sp = (sp + 127u) & ~127u;
  v.o_warp = sp;
  if (sp + 4u * (2048u + 128u + V3_RECCAP * 8u) > 227u * 1024u) return KEX_OK;
  // forward table: two bytes per lookup when the pair table fits 16-bit row offsets
  const bool pair = (size_t)NM * C * C * 2 <= 65534 && (C - 1) * C * 2 <= 255;
  const uint32_t rowlen = pair ? C * C : C;
  if ((size_t)NM * rowlen * 2 > 65534 || (C - 1) * 2 > 255) return KEX_OK;
  v.pair = pair ? 1u : 0u;
  v.rowbytes = rowlen * 2u;
  v.recip = (uint32_t)(((1ull << 32) + v.rowbytes - 1) / v.rowbytes);
  for (uint32_t r = 0; r < NM; ++r)
    if ((uint32_t)(((unsigned long long)(r * v.rowbytes) * v.recip) >> 32) != r) return KEX_OK;
  std::vector<uint16_t> fwd((size_t)NM * rowlen);
  for (uint32_t m = 0; m < NM; ++m)
    for (uint32_t c0 = 0; c0 < C; ++c0) {
      const uint32_t m1 = mulF[m * C + c0];
      if (!pair) { fwd[(size_t)m * C + c0] = (uint16_t)(m1 * v.rowbytes); continue; }
      for (uint32_t c1 = 0; c1 < C; ++c1) fwd[(size_t)m * rowlen + c0 * C + c1] = (uint16_t)(mulF[m1 * C + c1] * v.rowbytes);
    }
  v.fwd_entries = (uint32_t)fwd.size();
  ph.smem_fwd3 = 512 + ((fwd.size() * 2 + 15) & ~(size_t)15);
  ph.smem_seams3 = 256 + 4ull * Q1 * C + 2ull * NB * F.NG + 16;
  if (ph.smem_fwd3 > 200 * 1024 || ph.smem_seams3 > 200 * 1024) return KEX_OK;
  // element composition table: comp[a][b] = element of "a, then b" (the forward pass composes the
  // elements of the 128-byte blocks of a chunk across a warp)
  if (NM > 2048) return KEX_OK;
  std::vector<uint16_t> comp((size_t)NM * NM);
  {
    std::unordered_map<std::string, uint16_t> index;
    index.reserve(NM * 2);
    for (uint32_t m = 0; m < NM; ++m)
      index.emplace(std::string((const char *)(applyF + (size_t)m * Q1), Q1 * sizeof(uint16_t)), (uint16_t)m);
    std::vector<uint16_t> t(Q1);
    for (uint32_t a = 0; a < NM; ++a)
      for (uint32_t b = 0; b < NM; ++b) {
        for (uint32_t q = 0; q < Q1; ++q) t[q] = applyF[(size_t)b * Q1 + applyF[(size_t)a * Q1 + q]];
        auto it = index.find(std::string((const char *)t.data(), Q1 * sizeof(uint16_t)));
        if (it == index.end()) return KEX_ERR_BAD_BLOB;          // the element set is not closed
        comp[(size_t)a * NM + b] = it->second;
      }
  }
  v.NM = NM;
  const size_t o_be = 0, o_bw = o_be + 4ull * NE, o_tp = o_bw + 4ull * NE, o_fw = (o_tp + 4ull * NE + 15) & ~(size_t)15,
               o_cp = (o_fw + fwd.size() * 2 + 15) & ~(size_t)15, tot = o_cp + comp.size() * 2;
  std::vector<uint8_t> img(tot, 0);
  memcpy(img.data() + o_be, be3.data(), 4ull * NE);
  memcpy(img.data() + o_bw, be3w.data(), 4ull * NE);
  memcpy(img.data() + o_tp, tpl2.data(), 4ull * NE);
  memcpy(img.data() + o_fw, fwd.data(), fwd.size() * 2);
  memcpy(img.data() + o_cp, comp.data(), comp.size() * 2);
  CK(cudaMalloc(&ph.d_v3, tot));
  CK(cudaMemcpy(ph.d_v3, img.data(), tot, cudaMemcpyHostToDevice));
  v.be3 = (const uint32_t *)((const uint8_t *)ph.d_v3 + o_be);
  v.be3w = (const uint32_t *)((const uint8_t *)ph.d_v3 + o_bw);
  v.tpl2 = (const uint32_t *)((const uint8_t *)ph.d_v3 + o_tp);
  v.fwdtab = (const uint16_t *)((const uint8_t *)ph.d_v3 + o_fw);
  v.compF = (const uint16_t *)((const uint8_t *)ph.d_v3 + o_cp);
  v.ok = 1;
  if (getenv("KEX_DEBUG"))
    fprintf(stderr, "kexcuda: v3 kernels: entry stride %u B, %u emission entries, one-byte literals %s, forward table %s (%u entries), "
                    "%u forward elements, tables %u B\n", stride, NE, has_lit ? "direct" : "none", pair ? "pairs" : "single",
            v.fwd_entries, NM, v.o_warp);
  return KEX_OK;
}


// ---- G-mode tables (kex_v4.cuh).
// gtab for a choice of G: observed (state, live set) pairs first, most frequent
// first, each closed under the backward image (the live set of a predecessor
// follows from its successor's), then the blob's static choice for the states
// still open.  Same construction as fasttab.py build_gmode.
static void build_gtab(const PhaseHost &ph, const std::vector<std::pair<uint32_t, uint32_t>> &observed,
                       std::vector<uint8_t> &G, std::vector<uint32_t> &gtab) {
  const uint32_t Q = ph.dev.Q, Q1 = Q + 1, C = ph.dev.C, A = ph.dev.A, NL = ph.fdev.NL;
  const std::vector<uint32_t> &tr = ph.h_trans2, &BE = ph.h_BE;
  auto lam_before = [&](uint32_t l, uint32_t a) { return ((BE[l * A + a] & 0xFFFCu) / 4u) / A; };
  std::vector<std::vector<std::pair<uint32_t, uint32_t>>> preds(Q1);
  for (uint32_t q = 0; q < Q; ++q)
    for (uint32_t c = 0; c < C; ++c) {
      const uint32_t e = tr[q * C + c], q2 = e & 0xFFFFu;
      if (q2 != Q) preds[q2].push_back({q, (e >> 16) & 0xFFu});
    }
  G.assign(Q1, 0xFF);
  std::vector<uint32_t> work;
  auto close = [&](uint32_t q0) {
    work.assign(1, q0);
    while (!work.empty()) {
      const uint32_t q2 = work.back();
      work.pop_back();
      for (auto &pa : preds[q2])
        if (G[pa.first] == 0xFF) {
          G[pa.first] = (uint8_t)lam_before(G[q2], pa.second);
          work.push_back(pa.first);
        }
    }
  };
  for (auto &o : observed)
    if (o.first < Q && o.second < NL && G[o.first] == 0xFF) { G[o.first] = (uint8_t)o.second; close(o.first); }
  for (uint32_t q = 0; q < Q; ++q)
    if (G[q] == 0xFF && ph.g_static[q] != 0xFF) { G[q] = ph.g_static[q]; close(q); }
  gtab.assign((size_t)Q1 * C, Q);
  for (uint32_t q = 0; q < Q; ++q) {
    if (G[q] == 0xFF) continue;
    for (uint32_t c = 0; c < C; ++c) {
      const uint32_t e = tr[q * C + c], q2 = e & 0xFFFFu, a = (e >> 16) & 0xFFu;
      if (q2 == Q || G[q2] == 0xFF || lam_before(G[q2], a) != G[q]) continue;
      const uint32_t be = BE[(uint32_t)G[q2] * A + a], typ = be & 3u;
      uint32_t ln = (be >> 16) & 0xFFu, flags = 0;
      if (typ == 0u) ln = 0;
      else if (typ & 1u) { ln = 1; flags = V4_GB_C; }
      else flags = V4_GB_T | (be >> 24);
      gtab[q * C + c] = q2 | (ln << 16) | (flags << 24);
    }
  }
}

static int upload_gtab(kex_program *p, PhaseHost &ph, const std::vector<uint8_t> &G, const std::vector<uint32_t> &gtab,
                       cudaStream_t st) {
  CK(cudaMemcpyAsync((void *)ph.v4.gtab, gtab.data(), gtab.size() * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync((void *)ph.v4.G, G.data(), G.size(), cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));          // the vectors are the caller's
  return KEX_OK;
}

// The blob's G-mode sections (fasttab.py serialize_fast, header words 23..26).
static int load_v4(kex_program *p, PhaseHost &ph, const uint8_t *f, uint32_t fl, const uint16_t *applyF,
                   const uint32_t *trans2, const uint32_t *BE, const uint32_t *tplinfo, uint32_t NM, uint32_t NT) {
  V4Dev &v = ph.v4;
  memset(&v, 0, sizeof(v));
  if (!ph.v3.ok || getenv("KEX_NO_V4") || rd32(f, 92) != 1u) return KEX_OK;
  const FastDev &F = ph.fdev;
  const uint32_t Q1 = ph.dev.Q + 1, C = ph.dev.C, A = ph.dev.A, NL = F.NL;
  const uint32_t oG = rd32(f, 96), oT = rd32(f, 100), min_tpl = rd32(f, 104);
  if ((size_t)oG + Q1 > fl || (size_t)oT + 4ull * Q1 * C > fl) return KEX_ERR_BAD_BLOB;
  if (Q1 > 255 || NT > V4_MAX_TPL) return KEX_OK;
  const uint32_t *gt = (const uint32_t *)(f + oT);
  for (uint32_t i = 0; i < Q1 * C; ++i) {
    const uint32_t e = gt[i], fl8 = e >> 24;
    if ((e & 0xFFFFu) >= Q1 || ((fl8 & V4_GB_T) && (fl8 & 0x3Fu) >= NT) || ((fl8 & V4_GB_T) && (fl8 & V4_GB_C))) return KEX_ERR_BAD_BLOB;
  }
  for (uint32_t q = 0; q < Q1; ++q)
    if (f[oG + q] != 0xFF && f[oG + q] >= NL) return KEX_ERR_BAD_BLOB;
  // shared-memory layout: the replicated transition table must be addressable with 16 bits
  v.log = 7;
  size_t tbytes = (size_t)Q1 * (C + 1) * 128;
  // A table too large for per-lane copies stays on k3_emit: with two lanes per copy (LOG 6) the
  // G-mode kernel measured no faster there (iso_datetime: 19.6 vs 19.0 ms per 3.94 GiB).  The
  // variant is kept behind a knob for the tests.
  if (tbytes + 2048 > 65536 && !getenv("KEX_V4_LOG6")) return KEX_OK;
  if (getenv("KEX_V4_LOG6")) { v.log = 6; tbytes /= 2; }
  if (tbytes + 2048 > 65536) return KEX_OK;
  uint32_t sp = 0;
  v.o_trans = sp; sp += (uint32_t)tbytes;
  v.o_cls = sp; sp += 256;
  v.apply_smem = ((size_t)NM * Q1 <= 64 * 1024) ? 1u : 0u;
  v.o_apply = sp; if (v.apply_smem) sp += (NM * Q1 + 15u) & ~15u;
  v.o_tpl = sp; sp += V4_MAX_TPL * 4u;
  v.pool_stride = (F.pool_len + 4u + 12u + 3u) & ~3u;
  v.o_pool = sp; sp += 4u * v.pool_stride;
  sp = (sp + 15u) & ~15u;
  v.o_slots = sp; sp += 256u + 512u;
  sp = (sp + 127u) & ~127u;
  v.o_warp = sp;
  if (sp + 8u * (2048u + 128u + V4_RECCAP * 8u) > (uint32_t)V3_SMEM_MAX) return KEX_OK;
  v.rmw = (min_tpl >= 3u || NT == 0u) ? 1u : 0u;
  if (getenv("KEX_V4_NORMW")) v.rmw = 0u;              // knob: byte stores at the template edges (tests, timing)
  ph.h_trans2.assign(trans2, trans2 + (size_t)Q1 * C);
  ph.h_BE.assign(BE, BE + (size_t)NL * A);
  ph.g_static.assign(f + oG, f + oG + Q1);
  ph.g_static_tab.assign(gt, gt + (size_t)Q1 * C);
  std::vector<uint32_t> tpl(V4_MAX_TPL, 0);
  for (uint32_t t = 0; t < NT; ++t) {
    const uint32_t hm = tplinfo[2 * t + 1];
    uint32_t hole = 0;
    while (hm >> hole) ++hole;                              // hole offset + 1 (load_v3 has checked: at most one)
    tpl[t] = (tplinfo[2 * t] & 0xFFFFu) | (tplinfo[2 * t] & 0xFF0000u) | (hole << 24);
  }
  std::vector<uint8_t> ap8((size_t)NM * Q1);
  for (size_t i = 0; i < ap8.size(); ++i) ap8[i] = (uint8_t)applyF[i];
  auto al = [](size_t x) { return (x + 15) & ~(size_t)15; };
  const size_t o_gt = 0, o_g = al(o_gt + 4ull * Q1 * C), o_tp = al(o_g + Q1), o_ap = al(o_tp + 4ull * V4_MAX_TPL),
               tot = al(o_ap + ap8.size());
  std::vector<uint8_t> img(tot, 0);
  memcpy(img.data() + o_gt, gt, 4ull * Q1 * C);
  memcpy(img.data() + o_g, f + oG, Q1);
  memcpy(img.data() + o_tp, tpl.data(), 4ull * V4_MAX_TPL);
  memcpy(img.data() + o_ap, ap8.data(), ap8.size());
  CK(cudaMalloc(&ph.d_v4, tot));
  CK(cudaMemcpy(ph.d_v4, img.data(), tot, cudaMemcpyHostToDevice));
  const uint8_t *d = (const uint8_t *)ph.d_v4;
  v.gtab = (const uint32_t *)(d + o_gt);
  v.G = d + o_g;
  v.tpl = (const uint32_t *)(d + o_tp);
  v.apply8 = d + o_ap;
  v.ok = 1;
  ph.v4_learned = (NL == 1);                                // nothing to learn without registers
  if (getenv("KEX_DEBUG"))
    fprintf(stderr, "kexcuda: v4 (G-mode) emit kernel: tables %u B, element table %s, template edges %s\n", v.o_warp,
            v.apply_smem ? "in shared memory" : "in global memory", v.rmw ? "merged" : "byte stores");
  if (getenv("KEX_DEBUG") && v.log != 7) fprintf(stderr, "kexcuda: v4: entry stride %u B\n", 1u << v.log);
  return KEX_OK;
}

// Learn G from the exact live sets at the chunk boundaries of the shard that is
// about to be emitted (its state chain and live sets are on the device).
static int learn_gmode(kex_program *p, PhaseHost &ph, size_t ntiles, cudaStream_t st) {
  const size_t nchunks = (ntiles + V3_TPC - 1) / V3_TPC;
  size_t N = nchunks > 1 ? nchunks - 1 : 0;               // boundaries c = 1..N: state starts[c], live set lams[0][4c-1]
  if (N > 4096) N = 4096;
  if (N < 2) return KEX_OK;                               // too little to learn from; the static table stays
  std::vector<uint16_t> hs(N + 1);
  std::vector<uint8_t> hl(4 * N);
  CK(cudaMemcpyAsync(hs.data(), p->c->starts[0].p, (N + 1) * 2, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(hl.data(), p->c->lams[0].p, 4 * N, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  std::unordered_map<uint32_t, uint32_t> cnt;
  for (size_t c = 1; c <= N; ++c) cnt[(uint32_t)hs[c] << 8 | hl[4 * c - 1]]++;
  std::vector<std::pair<uint32_t, uint32_t>> order(cnt.begin(), cnt.end());
  std::sort(order.begin(), order.end(), [](const std::pair<uint32_t, uint32_t> &a, const std::pair<uint32_t, uint32_t> &b) {
    return a.second != b.second ? a.second > b.second : a.first < b.first;
  });
  std::vector<std::pair<uint32_t, uint32_t>> obs;
  for (auto &o : order) obs.push_back({o.first >> 8, o.first & 0xFFu});
  std::vector<uint8_t> G;
  std::vector<uint32_t> gtab;
  build_gtab(ph, obs, G, gtab);
  int rc = upload_gtab(p, ph, G, gtab, st);
  if (rc) return rc;
  ph.h_G = G;
  ph.v4_learned = true;
  if (getenv("KEX_DEBUG")) fprintf(stderr, "kexcuda: v4: G learnt from %zu chunk boundaries (%zu distinct pairs)\n", N, order.size());
  return KEX_OK;
}

// Fast section (fasttab.py): monoid tables.  Malformed -> KEX_ERR_BAD_BLOB;
// tables too large for shared memory -> the phase stays on the generic kernels.
static int load_fast(kex_program *p, const uint8_t *b, size_t len, PhaseHost &ph) {
  const uint32_t fo = rd32(b, 76), fl = rd32(b, 80);
  if (fo == 0 || getenv("KEX_FORCE_GENERIC")) return KEX_OK;
  if ((size_t)fo + fl > len || fl < 128) return KEX_ERR_BAD_BLOB;
  const uint8_t *f = b + fo;
  if (rd32(f, 0) != 0x4658454Bu || rd32(f, 4) != 1u || rd32(f, 40) != fl) return KEX_ERR_BAD_BLOB;
  const uint32_t Q1 = ph.dev.Q + 1, C = ph.dev.C, A = ph.dev.A;
  const uint32_t NM = rd32(f, 8), NL = rd32(f, 16), NG = rd32(f, 20), NB = rd32(f, 24), NT = rd32(f, 28),
                 pool_len = rd32(f, 32), max_emit = rd32(f, 36);
  if (rd32(f, 12) != Q1 || NM == 0 || NL == 0 || NL > 64 || NG == 0 || NG > 255 || NB == 0 || NB > 255 || NT > 255 ||
      A > 255 || C > 63 || (size_t)NM * C > 65535 || (size_t)NL * A * 4 > 65535)
    return KEX_ERR_BAD_BLOB;
  uint32_t off[12];
  for (int i = 0; i < 12; ++i) off[i] = rd32(f, 44 + 4 * i);
  const size_t need[12] = {2ull * NM * C, 2ull * NM * Q1, 4ull * Q1 * C, 4ull * NL * A, Q1,       (size_t)NB * NG,
                           (size_t)NB * NB, (size_t)NB * NL, NB,          8ull * NT,   pool_len, 4ull * NL};
  for (int i = 0; i < 12; ++i)
    if ((size_t)off[i] + need[i] > fl) return KEX_ERR_BAD_BLOB;
  const uint16_t *mulF = (const uint16_t *)(f + off[0]), *applyF = (const uint16_t *)(f + off[1]);
  const uint32_t *trans2 = (const uint32_t *)(f + off[2]), *BE = (const uint32_t *)(f + off[3]);
  for (size_t i = 0; i < (size_t)NM * C; ++i) if (mulF[i] >= NM) return KEX_ERR_BAD_BLOB;
  for (size_t i = 0; i < (size_t)NM * Q1; ++i) if (applyF[i] >= Q1) return KEX_ERR_BAD_BLOB;
  for (uint32_t i = 0; i < Q1 * C; ++i)
    if ((trans2[i] & 0xFFFFu) >= Q1 || ((trans2[i] >> 16) & 0xFFu) >= A || (trans2[i] >> 24) >= NG) return KEX_ERR_BAD_BLOB;
  const uint32_t *tplinfo = (const uint32_t *)(f + off[9]);
  for (uint32_t i = 0; i < NL * A; ++i) {
    const uint32_t e = BE[i];
    if ((e & 0xFFFCu) % (4 * A) != 0 || (e & 0xFFFCu) / (4 * A) >= NL) return KEX_ERR_BAD_BLOB;
    if ((e & 2u) && ((e >> 24) >= NT || (tplinfo[2 * (e >> 24)] >> 16) != ((e >> 16) & 0xFFu))) return KEX_ERR_BAD_BLOB;
  }
  for (uint32_t i = 0; i < NT; ++i)
    if ((tplinfo[2 * i] & 0xFFFFu) + (tplinfo[2 * i] >> 16) > pool_len) return KEX_ERR_BAD_BLOB;
  for (size_t i = 0; i < (size_t)NB * NG; ++i) if (f[off[5] + i] >= NB) return KEX_ERR_BAD_BLOB;
  for (size_t i = 0; i < (size_t)NB * NB; ++i) if (f[off[6] + i] >= NB) return KEX_ERR_BAD_BLOB;
  for (size_t i = 0; i < (size_t)NB * NL; ++i) if (f[off[7] + i] >= NL) return KEX_ERR_BAD_BLOB;
  ph.lam_final.assign(f + off[4], f + off[4] + Q1);
  for (uint8_t x : ph.lam_final) if (x != 0xFF && x >= NL) return KEX_ERR_BAD_BLOB;
  ph.lam_masks.assign((const uint32_t *)(f + off[11]), (const uint32_t *)(f + off[11]) + NL);

  auto al = [](size_t x) { return (x + 15) & ~(size_t)15; };
  // shared-memory layout of k_emit_fast: tables, then tile, staging window, records
  size_t tables = 4ull * Q1 * C + 4ull * NL * A + 8ull * NT + 2ull * NB * NG + 256 + (size_t)NB * NB + (size_t)NB * NL + 4 +
                  4ull * ((pool_len + 7) & ~3u);
  tables = (tables + 127) & ~(size_t)127;
  ph.smem_fm = 256 + al(2ull * NM * C);
  // limits of the kernels' packed fields; beyond them the generic kernels run the phase
  if (tables > 24 * 1024 || ph.smem_fm > 160 * 1024 || NG > 127 || (size_t)Q1 * C * 4 > 65535) return KEX_OK;
  const uint8_t *db = (const uint8_t *)ph.d_blob + fo;                  // the blob is already on the device
  FastDev &d = ph.fdev;
  d.NM = NM; d.NL = NL; d.NG = NG; d.NB = NB; d.NT = NT; d.pool_len = pool_len; d.max_emit = max_emit;
  d.tbl_bytes = (uint32_t)tables;
  d.mulF = (const uint16_t *)(db + off[0]);
  d.applyF = (const uint16_t *)(db + off[1]);
  d.trans2 = (const uint32_t *)(db + off[2]);
  d.BE = (const uint32_t *)(db + off[3]);
  d.lam_final = db + off[4];
  d.mulB = db + off[5];
  d.compB = db + off[6];
  d.applyB = db + off[7];
  d.constB = db + off[8];
  d.tplinfo = (const uint32_t *)(db + off[9]);
  d.pool = db + off[10];
  ph.smem_ef_tables = tables;
  ph.fast = true;
  int rc = load_v3(p, ph, mulF, applyF, BE, tplinfo, f + off[10], NM);
  if (rc) return rc;
  return load_v4(p, ph, f, fl, applyF, trans2, BE, tplinfo, NM, NT);
}

static int load_phase(kex_program *p, const uint8_t *b, size_t len, PhaseHost &ph) {
  if (len >= 32 && rd32(b, 0) == 0x4158454Bu) {          // "KEXA": action-interpreter phase
    if (rd32(b, 4) != 1u || rd32(b, 12) != ACT_ESC || rd32(b, 16) != len) return KEX_ERR_BAD_BLOB;
    const uint32_t nregs = rd32(b, 8);
    if (nregs + 1u > ACT_NSLOT) return KEX_ERR_UNSUPPORTED;
    memset(&ph.dev, 0, sizeof(ph.dev));
    ph.dev.R = nregs;
    ph.dev.max_out = 1;                                    // the interpreter never adds bytes
    ph.act = true;
    ph.act_nregs = nregs;
    return KEX_OK;
  }
  if (len < 96 || rd32(b, 0) != 0x5058454Bu || rd32(b, 4) != 1u) return KEX_ERR_BAD_BLOB;
  const uint32_t Q = rd32(b, 8), C = rd32(b, 12), R = rd32(b, 16), A = rd32(b, 20);
  const uint32_t npieces = rd32(b, 24), nconst = rd32(b, 28), init = rd32(b, 32), maxout = rd32(b, 36);
  uint32_t off[8];
  for (int i = 0; i < 8; ++i) off[i] = rd32(b, 40 + 4 * i);
  const uint32_t total = rd32(b, 72);
  if (total != len || R == 0 || R > 32 || Q >= 0xFFFF || A >= 0xFFFF || C == 0 || C > 256 || init >= Q)
    return KEX_ERR_BAD_BLOB;
  const uint32_t Q1 = Q + 1;
  if ((size_t)off[0] + 256 > len || (size_t)off[1] + 4ull * Q1 * C > len || (size_t)off[2] + 4ull * Q1 > len ||
      (size_t)off[3] + sizeof(ActHdr) * (size_t)A > len || (size_t)off[4] + (size_t)A * R > len ||
      (size_t)off[5] + 4ull * A * R > len || (size_t)off[6] + 8ull * npieces > len ||
      (size_t)off[7] + nconst > len)
    return KEX_ERR_BAD_BLOB;
  if ((size_t)Q1 * C >= 65536) return KEX_ERR_UNSUPPORTED;
  const uint32_t *trans = (const uint32_t *)(b + off[1]);
  ph.fin.assign((const int32_t *)(b + off[2]), (const int32_t *)(b + off[2]) + Q1);
  ph.acts.assign((const ActHdr *)(b + off[3]), (const ActHdr *)(b + off[3]) + A);
  ph.fate.assign(b + off[4], b + off[4] + (size_t)A * R);
  ph.pieces.assign((const Piece *)(b + off[6]), (const Piece *)(b + off[6]) + npieces);
  ph.consts.assign(b + off[7], b + off[7] + nconst);
  for (uint32_t i = 0; i < Q1 * C; ++i)
    if ((trans[i] & 0xFFFFu) > Q || (trans[i] >> 16) >= A) return KEX_ERR_BAD_BLOB;
  for (uint32_t a = 0; a < A; ++a) {
    if ((size_t)ph.acts[a].piece_off + ph.acts[a].npieces > npieces || ph.acts[a].kind > 3) return KEX_ERR_BAD_BLOB;
    if (ph.acts[a].outlen0 >= (1u << 24)) return KEX_ERR_UNSUPPORTED;
  }
  for (auto &pc : ph.pieces)
    if (pc.target >= R || (pc.kind == PIECE_CONST && (size_t)pc.off + pc.len > nconst)) return KEX_ERR_BAD_BLOB;
  for (int32_t f : ph.fin) if (f >= (int32_t)A) return KEX_ERR_BAD_BLOB;

  // derived tables
  std::vector<uint16_t> nextpm((size_t)Q1 * C);
  for (uint32_t i = 0; i < Q1 * C; ++i) nextpm[i] = (uint16_t)((trans[i] & 0xFFFFu) * C);
  std::vector<uint32_t> actinfo(A);
  for (uint32_t a = 0; a < A; ++a) actinfo[a] = ph.acts[a].kind | (ph.acts[a].outlen0 << 8);
  std::vector<uint32_t> img_cnt(C, 0);
  std::vector<uint16_t> img_state((size_t)C * Q, (uint16_t)(Q * C));
  std::vector<uint16_t> img_idx((size_t)C * Q1, 0xFFFF);
  for (uint32_t c = 0; c < C; ++c) {
    for (uint32_t q = 0; q < Q; ++q) {
      const uint32_t t = trans[q * C + c] & 0xFFFFu;
      if (t == Q) continue;
      uint32_t k = 0;
      for (; k < img_cnt[c]; ++k) if (img_state[(size_t)c * Q + k] == t * C) break;
      if (k == img_cnt[c]) img_state[(size_t)c * Q + img_cnt[c]++] = (uint16_t)(t * C);
      img_idx[(size_t)c * Q1 + q] = (uint16_t)k;
    }
  }
  CK(cudaMalloc(&ph.d_blob, len));
  CK(cudaMemcpy(ph.d_blob, b, len, cudaMemcpyHostToDevice));
  auto al = [](size_t x) { return (x + 15) & ~(size_t)15; };
  const size_t o_next = 0, o_info = al(o_next + nextpm.size() * 2), o_cnt = al(o_info + actinfo.size() * 4),
               o_ist = al(o_cnt + img_cnt.size() * 4), o_iidx = al(o_ist + img_state.size() * 2),
               tot = al(o_iidx + img_idx.size() * 2);
  std::vector<uint8_t> extra(tot, 0);
  memcpy(extra.data() + o_next, nextpm.data(), nextpm.size() * 2);
  memcpy(extra.data() + o_info, actinfo.data(), actinfo.size() * 4);
  memcpy(extra.data() + o_cnt, img_cnt.data(), img_cnt.size() * 4);
  memcpy(extra.data() + o_ist, img_state.data(), img_state.size() * 2);
  memcpy(extra.data() + o_iidx, img_idx.data(), img_idx.size() * 2);
  CK(cudaMalloc(&ph.d_extra, tot));
  CK(cudaMemcpy(ph.d_extra, extra.data(), tot, cudaMemcpyHostToDevice));
  const uint8_t *db = (const uint8_t *)ph.d_blob, *de = (const uint8_t *)ph.d_extra;
  PhaseDev &d = ph.dev;
  d.Q = Q; d.C = C; d.R = R; d.A = A; d.init = init; d.max_out = maxout;
  d.cls = db + off[0];
  d.trans = (const uint32_t *)(db + off[1]);
  d.acts = (const ActHdr *)(db + off[3]);
  d.fate = db + off[4];
  d.addlen = (const uint32_t *)(db + off[5]);
  d.pieces = (const Piece *)(db + off[6]);
  d.consts = db + off[7];
  d.nextpm = (const uint16_t *)(de + o_next);
  d.actinfo = (const uint32_t *)(de + o_info);
  d.img_cnt = (const uint32_t *)(de + o_cnt);
  d.img_state = (const uint16_t *)(de + o_ist);
  d.img_idx = (const uint16_t *)(de + o_iidx);
  ph.mask_bytes = (R == 1) ? 0 : (R <= 8 ? 1 : 4);
  ph.smem_maps = 256 + al((size_t)Q1 * C * 2);
  ph.smem_walk = 256 + (size_t)Q1 * C * 4 + (size_t)A * 4;
  ph.smem_emit = KEX_CHUNK + KEX_STAGE + 32 + (size_t)KEX_CHUNK * 2 + (size_t)KEX_CHUNK * ph.mask_bytes +
                 (ph.mask_bytes ? 2u * KEX_NT * 32u : 0u) + (size_t)Q1 * C * 4 + (size_t)A * 4 + 256;
  if (ph.smem_emit > 200 * 1024 || ph.smem_walk > 200 * 1024) return KEX_ERR_UNSUPPORTED;
  return load_fast(p, b, len, ph);
}

extern "C" int kex_load(const void *blob, size_t blob_len, int device, kex_program **out) {
  if (!blob || !out || blob_len < 16) return KEX_ERR_ARG;
  const uint8_t *b = (const uint8_t *)blob;
  if (rd32(b, 0) != 0x4C58454Bu || rd32(b, 4) != 1u) return KEX_ERR_BAD_BLOB;
  const uint32_t nph = rd32(b, 8);
  if (nph == 0 || nph > 64 || rd32(b, 12) != blob_len || 16 + 8ull * nph > blob_len) return KEX_ERR_BAD_BLOB;
  kex_program *p = new kex_program();
  p->device = device;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) { delete p; return KEX_ERR_CUDA; }
  p->phases.resize(nph);
  for (uint32_t i = 0; i < nph; ++i) {
    const uint32_t off = rd32(b, 16 + 8 * i), len = rd32(b, 20 + 8 * i);
    int rc = ((size_t)off + len > blob_len) ? KEX_ERR_BAD_BLOB : load_phase(p, b + off, len, p->phases[i]);
    if (rc != KEX_OK) { kex_free(p); return rc; }
  }
  // the attribute is per function, not per handle: always allow the device maximum
  const int mx = 200 * 1024;
  cudaFuncSetAttribute(k_chunk_maps, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_true_walk, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_emit<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_emit<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_emit<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_fwd_monoid, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_emit_fast<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k_emit_fast<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k3_fwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k3_fwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  cudaFuncSetAttribute(k3_seams, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
#define V3_EACH(X) X(7, true, true) X(7, true, false) X(7, false, true) X(7, false, false) \
                   X(6, true, true) X(6, true, false) X(6, false, true) X(6, false, false) \
                   X(5, true, true) X(5, true, false) X(5, false, true) X(5, false, false)
#define V3_ATTR(L, R, T) cudaFuncSetAttribute(k3_emit<L, R, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, V3_SMEM_MAX);
  V3_EACH(V3_ATTR)
#undef V3_ATTR
  cudaFuncSetAttribute(k4_emit<true, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, V3_SMEM_MAX);
  cudaFuncSetAttribute(k4_emit<false, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, V3_SMEM_MAX);
  cudaFuncSetAttribute(k4_emit<true, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, V3_SMEM_MAX);
  cudaFuncSetAttribute(k4_emit<false, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, V3_SMEM_MAX);
  cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, device);
  for (Ctx &c : p->cx) {
    if (cudaMallocHost((void **)&c.ctl_host, sizeof(FastCtl)) != cudaSuccess) { kex_free(p); return KEX_ERR_CUDA; }
    if (cudaMallocHost((void **)&c.res_host, sizeof(RunResult)) != cudaSuccess) { kex_free(p); return KEX_ERR_CUDA; }
    if (cudaHostAlloc((void **)&c.hpub, 4096, cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void **)&c.dpub, c.hpub, 0) != cudaSuccess) {
      cudaGetLastError();
      if (c.hpub) cudaFreeHost(c.hpub);
      c.hpub = c.dpub = nullptr;                       // fall back to copies
    }
  }
  for (int i = 0; i < 8; ++i) cudaEventCreate(&p->ev[i]);
  p->ev_ok = true;
  *out = p;
  return KEX_OK;
}

extern "C" void kex_free(kex_program *p) {
  if (!p) return;
  cudaSetDevice(p->device);
  for (auto &ph : p->phases) { cudaFree(ph.d_blob); cudaFree(ph.d_extra); cudaFree(ph.d_v3); cudaFree(ph.d_v4); }
  for (Ctx &c : p->cx) {
    for (int i = 0; i < 8; ++i) {
      cudaFree(c.maps[i].p); cudaFree(c.starts[i].p); cudaFree(c.fates[i].p); cudaFree(c.lives[i].p);
      cudaFree(c.bmaps[i].p); cudaFree(c.lams[i].p);
    }
    Buf *bs[] = {&c.samples, &c.blockpre, &c.pend, &c.resolved, &c.fail, &c.outlen, &c.outoff, &c.bsum, &c.res_dev, &c.desc, &c.ctl};
    for (Buf *b : bs) cudaFree(b->p);
    if (c.ctl_host) cudaFreeHost(c.ctl_host);
    if (c.res_host) cudaFreeHost(c.res_host);
    if (c.hpub) cudaFreeHost(c.hpub);
  }
  for (Buf *b : {&p->inter[0], &p->inter[1], &p->hostio_in, &p->hostio_out, &p->sin[0], &p->sin[1], &p->sout}) cudaFree(b->p);
  {
    ActScratch &a = p->as;
    Buf *abs[] = {&a.delta, &a.mn, &a.mx, &a.h0, &a.bsum, &a.fate, &a.add, &a.gfate, &a.gadd, &a.gvec, &a.vec, &a.wlen, &a.ctl};
    for (Buf *b : abs) cudaFree(b->p);
  }
  if (p->s_h2d) cudaStreamDestroy(p->s_h2d);
  if (p->s_comp) cudaStreamDestroy(p->s_comp);
  if (p->s_d2h) cudaStreamDestroy(p->s_d2h);
  for (cudaEvent_t e : p->pipe_ev) cudaEventDestroy(e);
  if (p->ev_ok) for (int i = 0; i < 8; ++i) cudaEventDestroy(p->ev[i]);
  delete p;
}

extern "C" int kex_info(const kex_program *p, uint32_t phase, kex_info_t *info) {
  if (!p || !info || phase >= p->phases.size()) return KEX_ERR_ARG;
  const PhaseDev &d = p->phases[phase].dev;
  info->nphases = (uint32_t)p->phases.size();
  info->nstates = d.Q; info->nclasses = d.C; info->nregs = d.R; info->nactions = d.A;
  info->max_out_per_byte = d.max_out; info->chunk_bytes = (uint32_t)p->phases[phase].tile();
  info->monoid_kernels = p->phases[phase].fast ? 1u : 0u;
  if (p->phases[phase].act) info->chunk_bytes = 1024u;
  const PhaseHost &ph = p->phases[phase];
  info->emit_kernel = (ph.v4.ok && !ph.v4_off) ? 4u : ph.v3.ok ? 3u : ph.fast ? 2u : 1u;
  info->exact_tiles = ph.v4_last_exact;
  return KEX_OK;
}

static uint32_t final_code(const PhaseHost &ph, uint32_t state) {
  const int32_t a = ph.fin[state];
  if (a < 0) return 0u;
  return ph.fast ? (uint32_t)ph.lam_final[state] : ph.acts[a].flush_mask;
}

extern "C" int kex_final_action(const kex_program *p, uint32_t state, int *accepting, uint32_t *seam_code,
                                const uint8_t **tail, size_t *tail_len) {
  if (!p) return KEX_ERR_ARG;
  const PhaseHost &ph = p->phases[p->c->sh_phase];
  if (ph.act) return KEX_ERR_UNSUPPORTED;
  if (state > ph.dev.Q) return KEX_ERR_ARG;
  const int32_t a = ph.fin[state];
  if (accepting) *accepting = a >= 0;
  if (seam_code) *seam_code = final_code(ph, state);
  // the tail is the concatenation of the action's constant pieces (all target the stream)
  static thread_local std::vector<uint8_t> buf;
  buf.clear();
  if (a >= 0)
    for (uint32_t k = 0; k < ph.acts[a].npieces; ++k) {
      const Piece &pc = ph.pieces[ph.acts[a].piece_off + k];
      buf.insert(buf.end(), ph.consts.begin() + pc.off, ph.consts.begin() + pc.off + pc.len);
    }
  if (tail) *tail = buf.data();
  if (tail_len) *tail_len = buf.size();
  return KEX_OK;
}

extern "C" size_t kex_out_bound(const kex_program *p, size_t n) {
  if (!p) return 0;
  size_t cur = n;
  for (auto &ph : p->phases) {
    size_t tail = 0;
    for (auto &a : ph.acts) if (a.total_len > tail) tail = a.total_len;
    cur = cur * ph.dev.max_out + tail;
  }
  return cur;
}

extern "C" int kex_select_phase(kex_program *p, uint32_t phase) {
  if (!p || phase > p->phases.size()) return KEX_ERR_ARG;
  p->only_phase = phase;
  return KEX_OK;
}
extern "C" uint32_t kex_last_launch_count(const kex_program *p) { return p ? p->launches : 0; }
extern "C" int kex_set_timing(kex_program *p, int enabled) { if (!p) return KEX_ERR_ARG; p->timing = enabled != 0; return KEX_OK; }
extern "C" float kex_last_kernel_ms(const kex_program *p, uint32_t which) { return (p && which < 4) ? p->ms[which] : 0.f; }
extern "C" const char *kex_last_cuda_error(const kex_program *p) { return p ? p->cuda_err.c_str() : ""; }
extern "C" const char *kex_strerror(int code) {
  switch (code) {
    case KEX_OK: return "ok";
    case KEX_ERR_BAD_BLOB: return "malformed kexprog blob";
    case KEX_ERR_CUDA: return "CUDA runtime error";
    case KEX_ERR_OUT_CAP: return "output buffer too small";
    case KEX_ERR_UNSUPPORTED: return "program exceeds a device-table limit";
    case KEX_ERR_ARG: return "bad argument";
  }
  return "unknown error";
}

// ---------------------------------------------------------------- shard steps
static int do_summarize_fast(kex_program *p, uint32_t phase, const uint8_t *d_in, size_t n, cudaStream_t st);
static int do_walk_fast(kex_program *p, uint32_t start_state, cudaStream_t st, size_t tail_first = 0);
static int do_emit_fast(kex_program *p, uint32_t lam_end, size_t n_eff, uint8_t *d_out, size_t out_cap, size_t *out_len,
                        cudaStream_t st);

static int do_summarize(kex_program *p, uint32_t phase, const uint8_t *d_in, size_t n, cudaStream_t st) {
  PhaseHost &ph = p->phases[phase];
  if (ph.fast) return do_summarize_fast(p, phase, d_in, n, st);
  const PhaseDev &P = ph.dev;
  const uint32_t Q1 = P.Q + 1;
  if (((uintptr_t)d_in & 15u) != 0) return KEX_ERR_ARG;
  p->c->sh_in = d_in; p->c->sh_n = n; p->c->sh_phase = phase;
  const size_t nchunks = (n + KEX_CHUNK - 1) / KEX_CHUNK;
  p->c->sh_nchunks = nchunks;
  // level sizes
  p->c->nlevels = 0;
  size_t c = nchunks;
  while (true) {
    p->c->lvl_count[p->c->nlevels++] = c;
    if (c <= 1) break;
    c = (c + KEX_FANIN - 1) / KEX_FANIN;
  }
  for (int l = 0; l < p->c->nlevels; ++l) {
    int rc = ensure(p, p->c->maps[l], p->c->lvl_count[l] * Q1 * sizeof(uint16_t));
    if (rc) return rc;
    rc = ensure(p, p->c->starts[l], p->c->lvl_count[l] * sizeof(uint16_t));
    if (rc) return rc;
  }
  int rc = ensure(p, p->c->res_dev, sizeof(RunResult));
  if (rc) return rc;
  if (p->timing) CK(cudaEventRecord(p->ev[0], st));
  k_chunk_maps<<<(unsigned)((nchunks + 127) / 128), 128, ph.smem_maps, st>>>(P, d_in, n, nchunks, (uint16_t *)p->c->maps[0].p);
  p->launches++;
  if (p->timing) CK(cudaEventRecord(p->ev[1], st));
  const unsigned bt = Q1 < 32 ? 32 : (Q1 > 256 ? 256 : ((Q1 + 31) / 32 * 32));
  for (int l = 1; l < p->c->nlevels; ++l) {
    k_compose<uint16_t, false><<<(unsigned)p->c->lvl_count[l], bt, 0, st>>>((const uint16_t *)p->c->maps[l - 1].p, p->c->lvl_count[l - 1],
                                                                       (uint16_t *)p->c->maps[l].p, Q1);
    p->launches++;
  }
  CK(cudaGetLastError());
  return KEX_OK;
}

static int do_walk(kex_program *p, uint32_t start_state, cudaStream_t st) {
  PhaseHost &ph = p->phases[p->c->sh_phase];
  if (ph.fast) return do_walk_fast(p, start_state, st);
  const PhaseDev &P = ph.dev;
  const uint32_t Q1 = P.Q + 1, R = P.R;
  const size_t nchunks = p->c->sh_nchunks, n = p->c->sh_n;
  const int top = p->c->nlevels - 1;
  k_set_u16<<<1, 1, 0, st>>>((uint16_t *)p->c->starts[top].p, start_state);
  p->launches++;
  for (int l = top; l >= 1; --l) {
    const size_t np = p->c->lvl_count[l];
    k_push_states<<<(unsigned)((np + 127) / 128), 128, 0, st>>>((const uint16_t *)p->c->maps[l - 1].p, p->c->lvl_count[l - 1],
                                                                (const uint16_t *)p->c->starts[l].p, np,
                                                                (uint16_t *)p->c->starts[l - 1].p, Q1);
    p->launches++;
  }
  int rc;
  const size_t nsamp = nchunks * KEX_NT;
  if ((rc = ensure(p, p->c->samples, nsamp * sizeof(uint16_t)))) return rc;
  if ((rc = ensure(p, p->c->resolved, nchunks * sizeof(unsigned long long)))) return rc;
  if ((rc = ensure(p, p->c->fail, nchunks * sizeof(uint32_t)))) return rc;
  if ((rc = ensure(p, p->c->pend, nchunks * R * sizeof(uint32_t)))) return rc;
  if ((rc = ensure(p, p->c->fates[0], nchunks * R))) return rc;
  if (p->timing) CK(cudaEventRecord(p->ev[2], st));
  k_true_walk<<<(unsigned)((nchunks + 127) / 128), 128, ph.smem_walk, st>>>(
      P, p->c->sh_in, n, nchunks, (const uint16_t *)p->c->starts[0].p, (uint16_t *)p->c->samples.p, (uint8_t *)p->c->fates[0].p,
      (uint32_t *)p->c->pend.p, (unsigned long long *)p->c->resolved.p, (uint32_t *)p->c->fail.p, (RunResult *)p->c->res_dev.p);
  p->launches++;
  if (p->timing) CK(cudaEventRecord(p->ev[3], st));
  k_reduce_fail<<<1, 1024, 0, st>>>((const uint32_t *)p->c->fail.p, nchunks, (RunResult *)p->c->res_dev.p);
  p->launches++;
  return fetch_sync(p, p->c->res_host, p->c->res_dev.p, sizeof(RunResult), st);
}

// fate maps of the first `nchunks_eff` chunks -> composed fate map (host, R bytes)
static int do_fate_up(kex_program *p, size_t nchunks_eff, uint8_t *h_fate, cudaStream_t st) {
  PhaseHost &ph = p->phases[p->c->sh_phase];
  const uint32_t R = ph.dev.R;
  if (R == 1 || nchunks_eff == 0) {
    for (uint32_t r = 0; r < R; ++r) h_fate[r] = (uint8_t)r;
    return KEX_OK;
  }
  size_t cnt[8];
  int nl = 0;
  size_t c = nchunks_eff;
  while (true) { cnt[nl++] = c; if (c <= 1) break; c = (c + KEX_FANIN - 1) / KEX_FANIN; }
  for (int l = 1; l < nl; ++l) {
    int rc = ensure(p, p->c->fates[l], cnt[l] * R);
    if (rc) return rc;
    k_compose<uint8_t, true><<<(unsigned)cnt[l], 32, 0, st>>>((const uint8_t *)p->c->fates[l - 1].p, cnt[l - 1],
                                                            (uint8_t *)p->c->fates[l].p, R);
    p->launches++;
  }
  CK(cudaMemcpyAsync(h_fate, p->c->fates[nl - 1].p, R, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return KEX_OK;
}

static int do_emit(kex_program *p, uint32_t live_end_mask, size_t n_eff, uint8_t *d_out, size_t out_cap,
                   size_t *out_len, cudaStream_t st) {
  PhaseHost &ph = p->phases[p->c->sh_phase];
  if (ph.fast) return do_emit_fast(p, live_end_mask, n_eff, d_out, out_cap, out_len, st);
  const PhaseDev &P = ph.dev;
  const uint32_t R = P.R;
  *out_len = 0;
  if (n_eff == 0) return KEX_OK;
  const size_t nchunks = (n_eff + KEX_CHUNK - 1) / KEX_CHUNK;
  int rc;
  size_t cnt[8];
  int nl = 0;
  size_t c = nchunks;
  while (true) { cnt[nl++] = c; if (c <= 1) break; c = (c + KEX_FANIN - 1) / KEX_FANIN; }
  if (R > 1) {
    for (int l = 0; l < nl; ++l) if ((rc = ensure(p, p->c->lives[l], cnt[l] * sizeof(uint32_t)))) return rc;
    for (int l = 1; l < nl; ++l) if ((rc = ensure(p, p->c->fates[l], cnt[l] * R))) return rc;
    // the up-sweep over exactly these chunks (a failing shard has fewer chunks than it walked)
    for (int l = 1; l < nl; ++l) {
      k_compose<uint8_t, true><<<(unsigned)cnt[l], 32, 0, st>>>((const uint8_t *)p->c->fates[l - 1].p, cnt[l - 1],
                                                              (uint8_t *)p->c->fates[l].p, R);
      p->launches++;
    }
    k_set_u32<<<1, 1, 0, st>>>((uint32_t *)p->c->lives[nl - 1].p, live_end_mask);
    p->launches++;
    for (int l = nl - 1; l >= 1; --l) {
      k_push_live<<<(unsigned)((cnt[l] + 127) / 128), 128, 0, st>>>((const uint8_t *)p->c->fates[l - 1].p, cnt[l - 1],
                                                                   (const uint32_t *)p->c->lives[l].p, cnt[l],
                                                                   (uint32_t *)p->c->lives[l - 1].p, R);
      p->launches++;
    }
  } else {
    if ((rc = ensure(p, p->c->lives[0], sizeof(uint32_t)))) return rc;
  }
  if ((rc = ensure(p, p->c->outlen, nchunks * 8))) return rc;
  if ((rc = ensure(p, p->c->outoff, (nchunks + 1) * 8))) return rc;
  const size_t nb = (nchunks + SCAN_TILE - 1) / SCAN_TILE;
  if ((rc = ensure(p, p->c->bsum, nb * 8))) return rc;
  k_outlen<<<(unsigned)((nchunks + 255) / 256), 256, 0, st>>>((const uint32_t *)p->c->pend.p, (const unsigned long long *)p->c->resolved.p,
                                                             (const uint32_t *)p->c->lives[0].p, nchunks, R,
                                                             (unsigned long long *)p->c->outlen.p);
  k_scan_sums<<<(unsigned)nb, SCAN_THREADS, 0, st>>>((const unsigned long long *)p->c->outlen.p, nchunks, (unsigned long long *)p->c->bsum.p);
  k_scan_top<<<1, SCAN_THREADS, 0, st>>>((unsigned long long *)p->c->bsum.p, nb, (RunResult *)p->c->res_dev.p);
  k_scan_apply<<<(unsigned)nb, SCAN_THREADS, 0, st>>>((const unsigned long long *)p->c->outlen.p, nchunks,
                                                     (const unsigned long long *)p->c->bsum.p, (unsigned long long *)p->c->outoff.p);
  p->launches += 4;
  CK(cudaMemcpyAsync(p->c->res_host, p->c->res_dev.p, sizeof(RunResult), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const size_t total = (size_t)p->c->res_host->total_out;
  *out_len = total;
  if (total > out_cap) return KEX_ERR_OUT_CAP;
  if (p->timing) CK(cudaEventRecord(p->ev[4], st));
  const unsigned grid = (unsigned)nchunks;
  if (ph.mask_bytes == 0)
    k_emit<0><<<grid, KEX_NT, ph.smem_emit, st>>>(P, p->c->sh_in, n_eff, (const uint16_t *)p->c->samples.p,
                                                  (const unsigned long long *)p->c->outoff.p, (const uint32_t *)p->c->lives[0].p, d_out);
  else if (ph.mask_bytes == 1)
    k_emit<1><<<grid, KEX_NT, ph.smem_emit, st>>>(P, p->c->sh_in, n_eff, (const uint16_t *)p->c->samples.p,
                                                  (const unsigned long long *)p->c->outoff.p, (const uint32_t *)p->c->lives[0].p, d_out);
  else
    k_emit<4><<<grid, KEX_NT, ph.smem_emit, st>>>(P, p->c->sh_in, n_eff, (const uint16_t *)p->c->samples.p,
                                                  (const unsigned long long *)p->c->outoff.p, (const uint32_t *)p->c->lives[0].p, d_out);
  p->launches++;
  if (p->timing) CK(cudaEventRecord(p->ev[5], st));
  CK(cudaGetLastError());
  return KEX_OK;
}


// ------------------------------------------------------- monoid-kernel steps
static void level_counts(size_t nchunks, size_t *cnt, int *nl) {
  int n = 0;
  size_t c = nchunks;
  while (true) { cnt[n++] = c; if (c <= 1) break; c = (c + KEX_FANIN - 1) / KEX_FANIN; }
  *nl = n;
}

static int do_summarize_fast(kex_program *p, uint32_t phase, const uint8_t *d_in, size_t n, cudaStream_t st) {
  PhaseHost &ph = p->phases[phase];
  const PhaseDev &P = ph.dev;
  const uint32_t Q1 = P.Q + 1;
  if (((uintptr_t)d_in & 15u) != 0) return KEX_ERR_ARG;
  p->c->sh_in = d_in; p->c->sh_n = n; p->c->sh_phase = phase;
  const size_t nchunks = (n + KEX_CHUNK - 1) / KEX_CHUNK;
  p->c->sh_nchunks = nchunks;
  level_counts(nchunks, p->c->lvl_count, &p->c->nlevels);
  int rc;
  for (int l = 0; l < p->c->nlevels; ++l) {
    if ((rc = ensure(p, p->c->maps[l], p->c->lvl_count[l] * Q1 * sizeof(uint16_t)))) return rc;
    if ((rc = ensure(p, p->c->starts[l], p->c->lvl_count[l] * sizeof(uint16_t)))) return rc;
  }
  if ((rc = ensure(p, p->c->res_dev, sizeof(RunResult)))) return rc;
  if ((rc = ensure(p, p->c->samples, nchunks * (ph.v3.ok ? V3_SPC : KEX_NT) * sizeof(uint16_t)))) return rc;
  if (p->timing) CK(cudaEventRecord(p->ev[0], st));
  if (ph.v3.ok) {
    // one warp per pair of chunks, grid-stride over resident CTAs (the table is staged once per CTA)
    if ((rc = ensure(p, p->c->blockpre, nchunks * V3_BPC * sizeof(uint16_t)))) return rc;
    const size_t npairs = (nchunks + 1) / 2;
    const unsigned ft = 512u;
    size_t ctas = (npairs + ft / 32 - 1) / (ft / 32);
    size_t per_sm = (size_t)V3_SMEM_MAX / (ph.smem_fwd3 + 1024);
    if (per_sm > 2u) per_sm = 2u;                  // __launch_bounds__ of k3_fwd
    const size_t resident = (size_t)p->num_sms * (per_sm < 1 ? 1 : per_sm);
    if (ctas > resident) ctas = resident;
    if (ph.v3.pair)
      k3_fwd<true><<<(unsigned)ctas, ft, ph.smem_fwd3, st>>>(P, ph.fdev, ph.v3, d_in, n, nchunks, (uint16_t *)p->c->samples.p,
                                                            (uint16_t *)p->c->blockpre.p, (uint16_t *)p->c->maps[0].p);
    else
      k3_fwd<false><<<(unsigned)ctas, ft, ph.smem_fwd3, st>>>(P, ph.fdev, ph.v3, d_in, n, nchunks, (uint16_t *)p->c->samples.p,
                                                             (uint16_t *)p->c->blockpre.p, (uint16_t *)p->c->maps[0].p);
  } else {
    k_fwd_monoid<<<(unsigned)((nchunks + 127) / 128), 128, ph.smem_fm, st>>>(P, ph.fdev, d_in, n, nchunks,
                                                                            (uint16_t *)p->c->samples.p, (uint16_t *)p->c->maps[0].p);
  }
  p->launches++;
  if (p->timing) CK(cudaEventRecord(p->ev[1], st));
  const unsigned bt = Q1 < 32 ? 32 : (Q1 > 256 ? 256 : ((Q1 + 31) / 32 * 32));
  for (int l = 1; l < p->c->nlevels; ++l) {
    k_compose<uint16_t, false><<<(unsigned)p->c->lvl_count[l], bt, 0, st>>>((const uint16_t *)p->c->maps[l - 1].p, p->c->lvl_count[l - 1],
                                                                       (uint16_t *)p->c->maps[l].p, Q1);
    p->launches++;
  }
  CK(cudaGetLastError());
  return KEX_OK;
}

static int do_walk_fast(kex_program *p, uint32_t start_state, cudaStream_t st, size_t tail_first) {
  PhaseHost &ph = p->phases[p->c->sh_phase];
  const PhaseDev &P = ph.dev;
  const uint32_t Q1 = P.Q + 1, NL = ph.fdev.NL;
  const size_t nchunks = p->c->sh_nchunks, n = p->c->sh_n;
  const int top = p->c->nlevels - 1;
  p->c->tail_first = tail_first;
  k_set_u16<<<1, 1, 0, st>>>((uint16_t *)p->c->starts[top].p, start_state);
  p->launches++;
  for (int l = top; l >= 1; --l) {
    const size_t np = p->c->lvl_count[l];
    k_push_states<<<(unsigned)((np + 127) / 128), 128, 0, st>>>((const uint16_t *)p->c->maps[l - 1].p, p->c->lvl_count[l - 1],
                                                                (const uint16_t *)p->c->starts[l].p, np,
                                                                (uint16_t *)p->c->starts[l - 1].p, Q1);
    p->launches++;
  }
  int rc;
  if (ph.v3.ok) {
    const size_t ntiles_all = (n + V3_TILE - 1) / V3_TILE;
    // tail evaluation: only the tiles from tail_first on (a chunk boundary) are walked; a failure
    // anywhere before shows as the FAIL state at their start (it is absorbing)
    const size_t t0 = tail_first, ntiles = ntiles_all - t0;
    if ((rc = ensure(p, p->c->bmaps[0], ntiles * NL))) return rc;
    CK(cudaMemsetAsync(p->c->res_dev.p, 0xFF, sizeof(RunResult), st));              // fail_pos = none (and no uninitialised padding reaches k_publish)
    if (p->timing) CK(cudaEventRecord(p->ev[2], st));
    k3_seams<<<(unsigned)((ntiles + 255) / 256), 256, ph.smem_seams3, st>>>(
        P, ph.fdev, p->c->sh_in + t0 * V3_TILE, n - t0 * V3_TILE, ntiles,
        (const uint16_t *)p->c->blockpre.p + t0 * (V3_TILE / V3_BLK), (const uint16_t *)p->c->starts[0].p + t0 / V3_TPC,
        (const uint16_t *)p->c->maps[0].p + (t0 / V3_TPC) * Q1, (uint8_t *)p->c->bmaps[0].p, (RunResult *)p->c->res_dev.p);
    p->launches++;
    if (p->timing) CK(cudaEventRecord(p->ev[3], st));
  } else {
  if ((rc = ensure(p, p->c->fail, nchunks * sizeof(uint32_t)))) return rc;
  if ((rc = ensure(p, p->c->bmaps[0], nchunks * NL))) return rc;
  if (p->timing) CK(cudaEventRecord(p->ev[2], st));
  k_seams<<<(unsigned)((nchunks + 127) / 128), 128, 0, st>>>(P, ph.fdev, p->c->sh_in, n, nchunks, (const uint16_t *)p->c->starts[0].p,
                                                            (const uint16_t *)p->c->maps[0].p, (uint8_t *)p->c->bmaps[0].p,
                                                            (uint32_t *)p->c->fail.p, (RunResult *)p->c->res_dev.p);
  p->launches++;
  if (p->timing) CK(cudaEventRecord(p->ev[3], st));
  k_reduce_fail<<<1, 1024, 0, st>>>((const uint32_t *)p->c->fail.p, nchunks, (RunResult *)p->c->res_dev.p);
  p->launches++;
  }
  return fetch_sync(p, p->c->res_host, p->c->res_dev.p, sizeof(RunResult), st);
}

// backward up-sweep over the first nchunks_eff chunk maps; returns the level count
static int lam_up(kex_program *p, size_t nchunks_eff, size_t *cnt, int *nl, cudaStream_t st) {
  const uint32_t NL = p->phases[p->c->sh_phase].fdev.NL;
  level_counts(nchunks_eff, cnt, nl);
  for (int l = 1; l < *nl; ++l) {
    int rc = ensure(p, p->c->bmaps[l], cnt[l] * NL);
    if (rc) return rc;
    k_compose_rev<<<(unsigned)cnt[l], 32, 0, st>>>((const uint8_t *)p->c->bmaps[l - 1].p, cnt[l - 1], (uint8_t *)p->c->bmaps[l].p, NL);
    p->launches++;
  }
  return KEX_OK;
}

static int do_emit_fast(kex_program *p, uint32_t lam_end, size_t n_eff, uint8_t *d_out, size_t out_cap, size_t *out_len,
                        cudaStream_t st) {
  PhaseHost &ph = p->phases[p->c->sh_phase];
  const PhaseDev &P = ph.dev;
  const uint32_t NL = ph.fdev.NL;
  *out_len = 0;
  if (n_eff == 0) return KEX_OK;
  if (lam_end >= NL || ((uintptr_t)d_out & 15u) != 0) return KEX_ERR_ARG;
  const size_t ntiles = (n_eff + ph.tile() - 1) / ph.tile();
  if (ntiles >= 0xFFFFFFFFull) return KEX_ERR_UNSUPPORTED;
  int rc;
  size_t cnt[8];
  int nl = 0;
  // tail evaluation: bmaps / lams hold the tiles [tail_first, ntiles) only (run_phase never shortens
  // n_eff in that mode)
  const size_t tail_first = p->c->tail_first;
  level_counts(ntiles - tail_first, cnt, &nl);
  if (NL > 1) {
    if ((rc = lam_up(p, ntiles - tail_first, cnt, &nl, st))) return rc;
    for (int l = 0; l < nl; ++l) if ((rc = ensure(p, p->c->lams[l], cnt[l]))) return rc;
    k_set_u8<<<1, 1, 0, st>>>((uint8_t *)p->c->lams[nl - 1].p, lam_end);
    p->launches++;
    for (int l = nl - 1; l >= 1; --l) {
      k_push_lam<<<(unsigned)((cnt[l] + 127) / 128), 128, 0, st>>>((const uint8_t *)p->c->bmaps[l - 1].p, cnt[l - 1],
                                                                  (const uint8_t *)p->c->lams[l].p, cnt[l],
                                                                  (uint8_t *)p->c->lams[l - 1].p, NL);
      p->launches++;
    }
  } else {
    if ((rc = ensure(p, p->c->lams[0], 16))) return rc;
  }
  if ((rc = ensure(p, p->c->desc, ntiles * 8))) return rc;
  if ((rc = ensure(p, p->c->ctl, sizeof(FastCtl)))) return rc;
  CK(cudaMemsetAsync(p->c->desc.p, 0, ntiles * 8, st));
  CK(cudaMemsetAsync(p->c->ctl.p, 0, sizeof(FastCtl), st));
  if (ph.v4.ok && !ph.v4_off) {
    // G-mode emit kernel (kex_v4.cuh): same launch geometry as k3_emit below
    const V4Dev &V = ph.v4;
    if (!ph.v4_learned && (rc = learn_gmode(p, ph, ntiles, st))) return rc;
    if (const char *e = getenv("KEX_V4_STAGE")) { const long x = atol(e); if (x >= 256) ph.v4_stage = (uint32_t)x & ~127u; }
    if (const char *e = getenv("KEX_V4_RECCAP")) { const long x = atol(e); if (x >= 8) ph.v4_reccap = (uint32_t)x; }
    uint32_t force_exact = getenv("KEX_V4_EXACT") ? 1u : 0u;
    if (const char *e = getenv("KEX_V4_KNOCK")) force_exact |= (uint32_t)atoi(e) & ~1u;     // timing experiments only (wrong output)
    if (tail_first && getenv("KEX_V4_TAIL_TEST")) force_exact |= 64u;                       // tests: pretend tile 0 broke the induction
    const uint32_t warp_bytes = (ph.v4_stage + 128u + ph.v4_reccap * 8u + 127u) & ~127u;
    uint32_t nwork = (uint32_t)(((size_t)V3_SMEM_MAX - V.o_warp) / warp_bytes);
    if (nwork > 31u) nwork = 31u;
    if (const char *e = getenv("KEX_V3_WORKERS")) { const uint32_t x = (uint32_t)atoi(e); if (x >= 1 && x < nwork) nwork = x; }
    if (nwork < 1u) return KEX_ERR_UNSUPPORTED;
    const size_t smem4 = (size_t)V.o_warp + (size_t)nwork * warp_bytes;
    const uint32_t nwarp = nwork + 1u;
    const bool regs = NL > 1;
    int occ = 0;
#define V4_EACH(X) X(true, 7) X(false, 7) X(true, 6) X(false, 6)
#define V4_OCC(R, L)                                                                                                 \
    if (regs == R && V.log == L) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k4_emit<R, L>, (int)(nwarp * 32u), smem4));
    V4_EACH(V4_OCC)
#undef V4_OCC
    if (occ < 1) return KEX_ERR_UNSUPPORTED;
    const size_t ngroups = (ntiles + nwork - 1) / nwork;
    size_t ctas = ngroups;
    if (ctas > (size_t)occ * (size_t)p->num_sms) ctas = (size_t)occ * (size_t)p->num_sms;
    if (p->timing) CK(cudaEventRecord(p->ev[4], st));
#define V4_LAUNCH(R, L)                                                                                              \
    if (regs == R && V.log == L)                                                                                     \
      k4_emit<R, L><<<(unsigned)ctas, nwarp * 32u, smem4, st>>>(                                                     \
          P, ph.fdev, V, p->c->sh_in, n_eff, (uint32_t)ntiles, (const uint16_t *)p->c->samples.p,                   \
          (const uint16_t *)p->c->blockpre.p, (const uint16_t *)p->c->starts[0].p,                                  \
          (const uint8_t *)p->c->lams[0].p - tail_first, (uint32_t)tail_first,                                      \
          (unsigned long long *)p->c->desc.p, (FastCtl *)p->c->ctl.p, d_out, out_cap,                                \
          (unsigned long long)p->emit_out_off, ph.v4_stage, warp_bytes, ph.v4_reccap, force_exact);
    V4_EACH(V4_LAUNCH)
#undef V4_LAUNCH
#undef V4_EACH
    p->launches++;
    if (p->timing) CK(cudaEventRecord(p->ev[5], st));
    CK(cudaGetLastError());
    if ((rc = fetch_sync(p, p->c->ctl_host, p->c->ctl.p, sizeof(FastCtl), st))) return rc;
    if (p->c->ctl_host->error == 1u) { p->cuda_err = "emit: chained scan timed out"; return KEX_ERR_CUDA; }
    if (p->c->ctl_host->error == 2u) return KEX_NEED_EXACT;      // tail evaluation: a tile before the tail needs exact live sets
    const size_t total = (size_t)p->c->ctl_host->total_out;
    *out_len = total;
    if (p->c->ctl_host->overflow || total + p->emit_out_off > out_cap) return KEX_ERR_OUT_CAP;
    // tiles that were evaluated exactly: when many, G no longer fits the data -- learn it again from
    // the next run's live sets; a program whose live sets are not state-determined (thousand_sep:
    // they follow the digit count) goes back to k3_emit for good
    const size_t slow = p->c->ctl_host->ticket;
    ph.v4_last_exact = (uint32_t)slow;
    if (!force_exact && regs && ntiles >= 64 && slow * 8 > ntiles) {
      if (++ph.v4_relearns > 1) ph.v4_off = true; else ph.v4_learned = false;
      if (getenv("KEX_DEBUG")) fprintf(stderr, "kexcuda: v4: %zu of %zu tiles evaluated exactly -> %s\n", slow, ntiles,
                                       ph.v4_off ? "back to k3_emit" : "G will be learnt again");
    }
    if (getenv("KEX_V4_STAGE") || getenv("KEX_V4_RECCAP")) return KEX_OK;
    // size the staging windows and record slots for the next run from what this one saw
    const double per_tile = (double)total / (double)ntiles;
    uint32_t want = (uint32_t)(per_tile * 1.25) + 256;
    want = (want + 255u) & ~255u;
    if (want < 2048u) want = 2048u;
    uint32_t wrec = (p->c->ctl_host->pad + p->c->ctl_host->pad / 4u + 32u + 63u) & ~63u;
    if (wrec < V4_RECCAP) wrec = V4_RECCAP;
    if (wrec > 1024u) wrec = 1024u;
    if (wrec > ph.v4_reccap || wrec + 128u < ph.v4_reccap) ph.v4_reccap = wrec;
    const uint32_t stage_max = (uint32_t)(((V3_SMEM_MAX - V.o_warp) / 8u - 128u - ph.v4_reccap * 8u) & ~255u);   // keep >= 8 warps
    if (want > stage_max) want = stage_max;
    if (want > ph.v4_stage || want + 1024u < ph.v4_stage) ph.v4_stage = want;
    return KEX_OK;
  }
  if (tail_first) return KEX_ERR_ARG;              // only the G-mode kernel evaluates tails
  if (ph.v3.ok) {
    // one CTA per SM: up to 31 worker warps + 1 scan warp (as many workers as the
    // staging windows leave room for); every CTA must be resident because groups
    // are assigned statically and chained in order
    const V3Dev &V = ph.v3;
    // test knobs: force the slow paths (tiny staging window / record slots) or exact live sets
    if (const char *e = getenv("KEX_V3_STAGE")) { const long x = atol(e); if (x >= 2048) ph.v3_stage = (uint32_t)x & ~255u; }
    if (const char *e = getenv("KEX_V3_RECCAP")) { const long x = atol(e); if (x >= 8) ph.v3_reccap = (uint32_t)x; }
    const uint32_t nospec = getenv("KEX_V3_NOSPEC") ? 1u : 0u;
    const uint32_t warp_bytes = (ph.v3_stage + 128u + ph.v3_reccap * 8u + 127u) & ~127u;
    uint32_t nwork = (uint32_t)(((size_t)V3_SMEM_MAX - V.o_warp) / warp_bytes);
    if (nwork > 31u) nwork = 31u;
    if (const char *e = getenv("KEX_V3_WORKERS")) { const uint32_t x = (uint32_t)atoi(e); if (x >= 1 && x < nwork) nwork = x; }
    if (nwork < 1u) return KEX_ERR_UNSUPPORTED;
    const size_t smem3 = (size_t)V.o_warp + (size_t)nwork * warp_bytes;
    const uint32_t nwarp = nwork + 1u;
    int occ = 0;
    const bool regs = NL > 1, lit = V.has_lit != 0;
#define V3_OCC(L, R, T)                                                                                              \
    if (V.log == L && regs == R && lit == T)                                                                        \
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k3_emit<L, R, T>, (int)(nwarp * 32u), smem3));
    V3_EACH(V3_OCC)
#undef V3_OCC
    if (occ < 1) return KEX_ERR_UNSUPPORTED;
    const size_t ngroups = (ntiles + nwork - 1) / nwork;
    size_t ctas = ngroups;
    if (ctas > (size_t)occ * (size_t)p->num_sms) ctas = (size_t)occ * (size_t)p->num_sms;
    if (p->timing) CK(cudaEventRecord(p->ev[4], st));
#define V3_LAUNCH(L, R, T)                                                                                          \
    if (V.log == L && regs == R && lit == T)                                                                        \
      k3_emit<L, R, T><<<(unsigned)ctas, nwarp * 32u, smem3, st>>>(                                                 \
          P, ph.fdev, V, p->c->sh_in, n_eff, (uint32_t)ntiles, (const uint16_t *)p->c->samples.p,                   \
          (const uint16_t *)p->c->blockpre.p, (const uint16_t *)p->c->starts[0].p, (const uint8_t *)p->c->lams[0].p,  \
          (unsigned long long *)p->c->desc.p, (FastCtl *)p->c->ctl.p, d_out, out_cap,                                \
          (unsigned long long)p->emit_out_off, ph.v3_stage, warp_bytes, ph.v3_reccap, nospec);
    V3_EACH(V3_LAUNCH)
#undef V3_LAUNCH
    p->launches++;
    if (p->timing) CK(cudaEventRecord(p->ev[5], st));
    CK(cudaGetLastError());
    if ((rc = fetch_sync(p, p->c->ctl_host, p->c->ctl.p, sizeof(FastCtl), st))) return rc;
    if (p->c->ctl_host->error) { p->cuda_err = "emit: chained scan timed out"; return KEX_ERR_CUDA; }
    const size_t total = (size_t)p->c->ctl_host->total_out;
    *out_len = total;
    if (p->c->ctl_host->overflow || total + p->emit_out_off > out_cap) return KEX_ERR_OUT_CAP;
    if (getenv("KEX_V3_STAGE") || getenv("KEX_V3_RECCAP")) return KEX_OK;
    // size the staging windows for the next run from the observed out/in ratio
    const double per_tile = (double)total / (double)ntiles;
    uint32_t want = (uint32_t)(per_tile * 1.25) + 256;
    want = (want + 255u) & ~255u;
    if (want < 2048u) want = 2048u;
    // ... and the record slots from the most records a tile had
    uint32_t wrec = (p->c->ctl_host->pad + p->c->ctl_host->pad / 4u + 32u + 63u) & ~63u;
    if (wrec < V3_RECCAP) wrec = V3_RECCAP;
    if (wrec > 1024u) wrec = 1024u;
    if (wrec > ph.v3_reccap || wrec + 128u < ph.v3_reccap) ph.v3_reccap = wrec;
    const uint32_t stage_max = (uint32_t)(((V3_SMEM_MAX - V.o_warp) / 8u - 128u - ph.v3_reccap * 8u) & ~255u);   // keep >= 8 warps
    if (want > stage_max) want = stage_max;
    if (want > ph.v3_stage || want + 1024u < ph.v3_stage) ph.v3_stage = want;
    return KEX_OK;
  }
  const size_t smem = ph.smem_ef_tables + KEX_CHUNK + ph.stage_bytes + 32 + EF_RECCAP * 4 + 16;
  if (ph.ef_ctas_per_sm == 0 || ph.ef_stage_cfg != ph.stage_bytes) {
    int occ = 0;
    if (NL > 1) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_emit_fast<true>, EF_NT, smem));
    else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_emit_fast<false>, EF_NT, smem));
    if (occ < 1) return KEX_ERR_UNSUPPORTED;
    ph.ef_ctas_per_sm = occ;
    ph.ef_stage_cfg = ph.stage_bytes;
  }
  const size_t resident = (size_t)ph.ef_ctas_per_sm * p->num_sms;
  const unsigned grid = (unsigned)(ntiles < resident ? ntiles : resident);
  if (p->timing) CK(cudaEventRecord(p->ev[4], st));
  if (NL > 1)
    k_emit_fast<true><<<grid, EF_NT, smem, st>>>(P, ph.fdev, p->c->sh_in, n_eff, (uint32_t)ntiles, (const uint16_t *)p->c->samples.p,
                                                 (const uint16_t *)p->c->starts[0].p, (const uint8_t *)p->c->lams[0].p,
                                                 (unsigned long long *)p->c->desc.p, (FastCtl *)p->c->ctl.p, d_out, out_cap,
                                                 ph.stage_bytes);
  else
    k_emit_fast<false><<<grid, EF_NT, smem, st>>>(P, ph.fdev, p->c->sh_in, n_eff, (uint32_t)ntiles, (const uint16_t *)p->c->samples.p,
                                                  (const uint16_t *)p->c->starts[0].p, (const uint8_t *)p->c->lams[0].p,
                                                  (unsigned long long *)p->c->desc.p, (FastCtl *)p->c->ctl.p, d_out, out_cap,
                                                  ph.stage_bytes);
  p->launches++;
  if (p->timing) CK(cudaEventRecord(p->ev[5], st));
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(p->c->ctl_host, p->c->ctl.p, sizeof(FastCtl), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (p->c->ctl_host->error) { p->cuda_err = "emit: chained scan timed out"; return KEX_ERR_CUDA; }
  const size_t total = (size_t)p->c->ctl_host->total_out;
  *out_len = total;
  if (p->c->ctl_host->overflow || total > out_cap) return KEX_ERR_OUT_CAP;
  // size the staging window for the next run from the observed out/in ratio
  const double per_tile = (double)total / (double)ntiles;
  uint32_t want = (uint32_t)(per_tile * 1.25) + 512;
  want = (want + 1023u) & ~1023u;
  if (want < 8192u) want = 8192u;
  const uint32_t stage_max = (uint32_t)((65535 - 64 - KEX_CHUNK - ph.smem_ef_tables) & ~(size_t)1023);   // record offsets are 16 bits
  if (want > stage_max) want = stage_max;
  if (want > ph.stage_bytes || want + 4096u < ph.stage_bytes) ph.stage_bytes = want;
  return KEX_OK;
}

extern "C" int kex_shard_summarize(kex_program *p, const uint8_t *d_in, size_t n, uint16_t *h_state_map, void *stream) {
  if (!p || !h_state_map || (n && !d_in)) return KEX_ERR_ARG;
  if (p->phases.size() != 1) return KEX_ERR_UNSUPPORTED;
  CK(cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;
  p->launches = 0;
  const uint32_t Q1 = p->phases[0].dev.Q + 1;
  if (n == 0) {
    p->c->sh_in = d_in; p->c->sh_n = 0; p->c->sh_nchunks = 0; p->c->sh_phase = 0;
    for (uint32_t q = 0; q < Q1; ++q) h_state_map[q] = (uint16_t)q;
    return KEX_OK;
  }
  int rc = do_summarize(p, 0, d_in, n, st);
  if (rc) return rc;
  CK(cudaMemcpyAsync(h_state_map, p->c->maps[p->c->nlevels - 1].p, Q1 * sizeof(uint16_t), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return KEX_OK;
}

extern "C" size_t kex_seam_bytes(const kex_program *p) {
  if (!p || p->phases.empty()) return 0;
  const PhaseHost &ph = p->phases[0];
  return ph.fast ? ph.fdev.NL : ph.dev.R;
}

// Walk the summarized shard from its true start state: end state, first failing
// position, and the shard's seam summary (how the seam code at its end maps to
// its start).
static int shard_walk_seam(kex_program *p, uint32_t start_state, uint32_t *end_state, size_t *fail_pos, uint8_t *h_seam,
                           cudaStream_t st, bool allow_tail = false) {
  PhaseHost &ph = p->phases[p->c->sh_phase];
  const PhaseDev &P = ph.dev;
  if (start_state > P.Q) return KEX_ERR_ARG;
  const uint32_t nseam = ph.fast ? ph.fdev.NL : P.R;
  if (p->c->sh_n == 0) {
    *end_state = start_state; *fail_pos = (size_t)-1;
    for (uint32_t r = 0; r < nseam; ++r) h_seam[r] = (uint8_t)r;
    return KEX_OK;
  }
  int rc;
  // G-mode tail evaluation of a shard (see run_phase): k3_seams and the live-set tree over the last
  // tiles only.  The shard's seam summary -- how the live set at its end maps to its start -- is
  // then the constant map to G[start state], provided the tail's own summary is a constant map
  // whose value is G[state at the tail's start] (the anchor) and every tile before the tail is
  // G-consistent; the latter is only known after the emit, which reports KEX_RETRY_EXACT if not.
  const bool regs = ph.fast && ph.fdev.NL > 1;
  if (allow_tail && ph.v4.ok && !ph.v4_off && ph.v4_learned && ph.v4_tail_ok && start_state != P.Q &&
      (!regs || (ph.h_G.size() == (size_t)P.Q + 1 && ph.h_G[start_state] != 0xFF)) && !getenv("KEX_V4_NOTAIL") &&
      !getenv("KEX_V4_EXACT")) {
    const size_t ntiles_all = (p->c->sh_n + V3_TILE - 1) / V3_TILE;
    if (ntiles_all >= 4 * V4_TAIL_TILES) {
      const size_t tail_first = (ntiles_all - V4_TAIL_TILES) / V3_TPC * V3_TPC;
      if ((rc = do_walk_fast(p, start_state, st, tail_first))) return rc;
      if (p->c->res_host->fail_pos == KEX_NONE64 && p->c->res_host->end_state != P.Q && !regs) {
        // no registers: nothing to stitch, the tail walk was only needed to see a failure
        *fail_pos = (size_t)-1;
        *end_state = p->c->res_host->end_state;
        for (uint32_t r = 0; r < nseam; ++r) h_seam[r] = (uint8_t)r;
        return KEX_OK;
      }
      if (p->c->res_host->fail_pos == KEX_NONE64 && p->c->res_host->end_state != P.Q) {
        size_t cnt[8];
        int nl = 0;
        if ((rc = lam_up(p, ntiles_all - tail_first, cnt, &nl, st))) return rc;
        std::vector<uint8_t> bt(nseam);
        uint16_t s0 = 0;
        if ((rc = fetch_sync(p, bt.data(), p->c->bmaps[nl - 1].p, nseam, st))) return rc;
        if ((rc = fetch_sync(p, &s0, (const uint16_t *)p->c->starts[0].p + tail_first / V3_TPC, sizeof(s0), st))) return rc;
        bool constant = true;
        for (uint32_t r = 1; r < nseam; ++r) constant = constant && bt[r] == bt[0];
        if (constant && s0 <= P.Q && ph.h_G[s0] == bt[0]) {
          *fail_pos = (size_t)-1;
          *end_state = p->c->res_host->end_state;
          for (uint32_t r = 0; r < nseam; ++r) h_seam[r] = ph.h_G[start_state];
          return KEX_OK;
        }
      }
    }
  }
  if ((rc = do_walk(p, start_state, st))) return rc;
  const unsigned long long f = p->c->res_host->fail_pos;
  *fail_pos = (f == KEX_NONE64) ? (size_t)-1 : (size_t)f;
  *end_state = p->c->res_host->end_state;
  const size_t n_eff = (f == KEX_NONE64) ? p->c->sh_n : (size_t)f;
  const size_t unit = ph.fast ? ph.tile() : (size_t)KEX_CHUNK;
  const size_t nchunks_eff = (n_eff + unit - 1) / unit;
  if (!ph.fast) return do_fate_up(p, nchunks_eff, h_seam, st);
  if (nseam == 1 || nchunks_eff == 0) {
    for (uint32_t r = 0; r < nseam; ++r) h_seam[r] = (uint8_t)r;
    return KEX_OK;
  }
  size_t cnt[8];
  int nl = 0;
  if ((rc = lam_up(p, nchunks_eff, cnt, &nl, st))) return rc;
  return fetch_sync(p, h_seam, p->c->bmaps[nl - 1].p, nseam, st);
}

extern "C" int kex_shard_walk(kex_program *p, uint32_t start_state, uint32_t *end_state, size_t *fail_pos,
                              uint8_t *h_seam, void *stream) {
  if (!p || !end_state || !fail_pos || !h_seam) return KEX_ERR_ARG;
  CK(cudaSetDevice(p->device));
  return shard_walk_seam(p, start_state, end_state, fail_pos, h_seam, (cudaStream_t)stream, p->shard_tail);
}

extern "C" int kex_set_shard_tail(kex_program *p, int enabled) {
  if (!p) return KEX_ERR_ARG;
  p->shard_tail = enabled != 0;
  return KEX_OK;
}

// Seam codes at the end of every shard from the shards' seam summaries (in
// shard order) and the code of the end-of-input action.
extern "C" int kex_stitch_live(const kex_program *p, const uint8_t *seams, size_t nshards, uint32_t final_code,
                               uint32_t *codes) {
  if (!p || !seams || !codes) return KEX_ERR_ARG;
  const PhaseHost &ph = p->phases[0];
  const size_t nb = ph.fast ? ph.fdev.NL : ph.dev.R;
  uint32_t code = final_code;
  for (size_t r = nshards; r-- > 0;) {
    codes[r] = code;
    const uint8_t *m = seams + r * nb;
    if (ph.fast) {
      if (code >= nb) return KEX_ERR_ARG;
      code = m[code];
    } else {
      const uint32_t live = code | 1u;
      uint32_t nx = 0;
      for (uint32_t i = 1; i < nb; ++i)
        if (m[i] != KEX_DEAD && ((live >> m[i]) & 1u)) nx |= 1u << i;
      code = nx;
    }
  }
  return KEX_OK;
}

extern "C" int kex_shard_emit(kex_program *p, uint32_t live_end_mask, size_t n_eff, uint8_t *d_out, size_t out_cap,
                              size_t *out_len, void *stream) {
  if (!p || !out_len || n_eff > p->c->sh_n) return KEX_ERR_ARG;
  CK(cudaSetDevice(p->device));
  int rc = do_emit(p, live_end_mask, n_eff, d_out, out_cap, out_len, (cudaStream_t)stream);
  if (rc == KEX_NEED_EXACT) {
    // only reachable after kex_set_shard_tail: the shard's seam summary was given under an assumption
    // that did not hold; this program evaluates exactly from now on, the caller repeats the shard steps
    p->phases[p->c->sh_phase].v4_tail_ok = false;
    *out_len = 0;
    return KEX_RETRY_EXACT;
  }
  if (rc == KEX_OK && p->timing && p->ev_ok && cudaStreamSynchronize((cudaStream_t)stream) == cudaSuccess) {
    // kernel times of this shard's three steps (kex_last_kernel_ms): forward, seams, emit
    if (cudaEventElapsedTime(&p->ms[0], p->ev[0], p->ev[1]) != cudaSuccess) p->ms[0] = 0.f;
    if (cudaEventElapsedTime(&p->ms[1], p->ev[2], p->ev[3]) != cudaSuccess) p->ms[1] = 0.f;
    if (!*out_len || cudaEventElapsedTime(&p->ms[2], p->ev[4], p->ev[5]) != cudaSuccess) p->ms[2] = 0.f;
    p->ms[3] = 0.f;
    cudaGetLastError();
  }
  return rc;
}

// ------------------------------------------------------------ whole pipeline
// Action-interpreter phase over an action stream resident on the device
// (kex_act.cuh).  The stream of a stage that rejected in its transducer phase is
// interpreted as far as it goes (the bottom builder is the result).
static int run_phase_act(kex_program *p, uint32_t phase, const uint8_t *d_in, size_t n, uint8_t *d_out,
                         size_t out_cap, size_t *out_len, cudaStream_t st) {
  const PhaseHost &ph = p->phases[phase];
  *out_len = 0;
  if (n == 0) return KEX_OK;
  if (n >= 0xFFFFFFF0ull) return KEX_ERR_UNSUPPORTED;      // 32-bit positions
  uint32_t tile = 1024;
  if (const char *e = getenv("KEX_ACT_TILE")) { const long x = atol(e); if (x > 0) tile = (uint32_t)x; }
  tile = (tile + 15u) & ~15u;                              // tiles start on 16-byte boundaries (vector loads)
  if (((uintptr_t)d_in & 15u) || ((uintptr_t)d_out & 3u)) return KEX_ERR_ARG;
  const size_t ntiles = (n + tile - 1) / tile, ngroups = (ntiles + ACT_GROUP - 1) / ACT_GROUP;
  ActScratch &a = p->as;
  int rc;
  if ((rc = ensure(p, a.delta, 4 * ntiles)) || (rc = ensure(p, a.mn, 4 * ntiles)) || (rc = ensure(p, a.mx, 4 * ntiles)) ||
      (rc = ensure(p, a.h0, 4 * (ntiles + 1))) || (rc = ensure(p, a.bsum, 8 * ((ntiles + ACT_NT - 1) / ACT_NT + 1))) || (rc = ensure(p, a.fate, (size_t)ACT_NSLOT * ntiles)) ||
      (rc = ensure(p, a.add, 4ull * ACT_NSLOT * ntiles)) || (rc = ensure(p, a.gfate, (size_t)ACT_NSLOT * ngroups)) ||
      (rc = ensure(p, a.gadd, 4ull * ACT_NSLOT * ngroups)) || (rc = ensure(p, a.gvec, 4ull * ACT_NSLOT * (ngroups + 1))) ||
      (rc = ensure(p, a.vec, 4ull * ACT_NSLOT * ntiles)) || (rc = ensure(p, a.wlen, 4ull * (n / 2 + 1))) ||
      (rc = ensure(p, a.ctl, sizeof(ActCtl))))
    return rc;
  ActCtl *ctl = (ActCtl *)a.ctl.p;
  int32_t *delta = (int32_t *)a.delta.p, *mn = (int32_t *)a.mn.p, *mx = (int32_t *)a.mx.p, *h0 = (int32_t *)a.h0.p;
  uint8_t *fate = (uint8_t *)a.fate.p, *gfate = (uint8_t *)a.gfate.p;
  uint32_t *add = (uint32_t *)a.add.p, *gadd = (uint32_t *)a.gadd.p, *gvec = (uint32_t *)a.gvec.p,
           *vec = (uint32_t *)a.vec.p, *wlen = (uint32_t *)a.wlen.p;
  const unsigned tb = (unsigned)((ntiles + ACT_NT - 1) / ACT_NT), gb = (unsigned)((ngroups + 3) / 4);   // thread per tile, warp per group
  ActCtl h;
  CK(cudaMemsetAsync(ctl, 0, sizeof(ActCtl), st));
  long long *bsum = (long long *)a.bsum.p;
  ka_heights<<<tb, ACT_NT, 0, st>>>(d_in, n, tile, ntiles, ph.act_nregs, delta, mn, mx, bsum, ctl);
  ka_height_scan<<<1, 1024, 0, st>>>(bsum, tb, ctl);
  ka_tile_heights<<<tb, ACT_NT, 0, st>>>(delta, mn, mx, bsum, ntiles, ph.act_nregs, h0, ctl);
  p->launches += 3;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(&h, ctl, sizeof(h), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (h.err & 5u) return KEX_ERR_ARG;                      // not an action stream
  if (h.err & 2u) return KEX_ERR_UNSUPPORTED;              // builders nested deeper than the slots allow
  // slots in use: builders 0..hmax and the registers (checked above to fit the 32 lanes of the scan kernels)
  const uint32_t nslots = (uint32_t)h.hmax + 1u + ph.act_nregs;
  const size_t cols = (size_t)nslots * ACT_NT;
  ka_fwd_summary<<<tb, ACT_NT, cols * 8, st>>>(d_in, n, tile, ntiles, h0, nslots, fate, add);
  ka_group_compose<<<gb, 128, 0, st>>>(fate, add, ntiles, ngroups, gfate, gadd);
  ka_group_scan<<<1, 32, 0, st>>>(gfate, gadd, ngroups, gvec);
  ka_tile_vectors<<<gb, 128, 0, st>>>(fate, add, ntiles, ngroups, gvec, vec);
  ka_fwd_exact<<<tb, ACT_NT, cols * 12, st>>>(d_in, n, tile, ntiles, h0, nslots, vec, wlen, fate, add, ctl);
  p->launches += 5;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(&h, ctl, sizeof(h), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  *out_len = h.total;
  if (h.total > out_cap) return KEX_ERR_OUT_CAP;
  if (h.total == 0) return KEX_OK;
  ka_group_bcompose<<<gb, 128, 0, st>>>(fate, add, ntiles, ngroups, gfate, gadd);
  ka_group_bscan<<<1, 32, 0, st>>>(gfate, gadd, ngroups, ctl, gvec);
  ka_tile_bvectors<<<gb, 128, 0, st>>>(fate, add, ntiles, ngroups, gvec, vec);
  ka_write<<<tb, ACT_NT, cols * 4, st>>>(d_in, n, tile, ntiles, h0, nslots, vec, wlen, d_out);
  p->launches += 4;
  CK(cudaGetLastError());
  return KEX_OK;
}

static int run_phase(kex_program *p, uint32_t phase, const uint8_t *d_in, size_t n, uint8_t *d_out, size_t out_cap,
                     size_t *out_len, int *status, size_t *fail_count, cudaStream_t st, bool feeds_interpreter = false) {
  PhaseHost &ph = p->phases[phase];
  if (ph.act) {
    // status and count are the transducer phase's (kex_run_device keeps them)
    *status = KEX_ACCEPT;
    *fail_count = 0;
    return run_phase_act(p, phase, d_in, n, d_out, out_cap, out_len, st);
  }
  const PhaseDev &P = ph.dev;
  p->c->sh_phase = phase;
  uint32_t end_state = P.init;
  size_t n_eff = n;
  bool failed = false;
  // G-mode tail evaluation (kex_v4.cuh): once G is learnt, the live set at every position of a
  // G-consistent run is G[state] -- by induction backwards from any tile boundary where the exact
  // live set equals G[state].  So k3_seams and the live-set tree only run over the last
  // V4_TAIL_TILES tiles (where the end of the input makes the live sets differ from G); the first
  // tail tile is the anchor: it must verify in G-mode.  Failures show as the FAIL state at the
  // tail's start.  Anything unexpected (a failure, an inconsistent tile before the tail) repeats
  // the phase with exact live sets everywhere and switches the shortcut off for this program.
  const size_t ntiles_all = (n + V3_TILE - 1) / V3_TILE;
  size_t tail_first = 0;
  if (ph.v4.ok && !ph.v4_off && ph.v4_learned && ph.v4_tail_ok && ntiles_all >= 4 * V4_TAIL_TILES && !getenv("KEX_V4_NOTAIL") &&
      !getenv("KEX_V4_EXACT"))
    tail_first = (ntiles_all - V4_TAIL_TILES) / V3_TPC * V3_TPC;
  if (n > 0) {
    int rc = do_summarize(p, phase, d_in, n, st);
    if (rc) return rc;
    if (tail_first) {
      if ((rc = do_walk_fast(p, P.init, st, tail_first))) return rc;
      if (p->c->res_host->fail_pos != KEX_NONE64 || p->c->res_host->end_state == P.Q) tail_first = 0;   // a failure: exact path
    }
    if (!tail_first && (rc = do_walk(p, P.init, st))) return rc;
    const unsigned long long f = p->c->res_host->fail_pos;
    if (f != KEX_NONE64) { failed = true; n_eff = (size_t)f; }
    end_state = p->c->res_host->end_state;
  } else {
    p->c->sh_in = d_in; p->c->sh_n = 0; p->c->sh_nchunks = 0;
  }
  const int32_t fa = failed ? -1 : ph.fin[end_state];
  const bool accept = fa >= 0;
  const uint32_t live = accept ? final_code(ph, end_state) : 0u;
  size_t body = 0;
  int rc = do_emit(p, live, n_eff, d_out, out_cap, &body, st);
  if (rc == KEX_NEED_EXACT) {
    ph.v4_tail_ok = false;
    if (getenv("KEX_DEBUG")) fprintf(stderr, "kexcuda: v4: tail evaluation met a tile that needs exact live sets -> exact run\n");
    if ((rc = do_walk(p, P.init, st))) return rc;
    rc = do_emit(p, live, n_eff, d_out, out_cap, &body, st);
  }
  if (rc == KEX_ERR_OUT_CAP) { *out_len = body + (accept ? ph.acts[fa].total_len : 0); return rc; }
  if (rc) return rc;
  if (accept) {
    const ActHdr &h = ph.acts[fa];
    size_t o = body;
    size_t tail = 0;
    for (uint32_t k = 0; k < h.npieces; ++k) tail += ph.pieces[h.piece_off + k].len;
    if (body + tail > out_cap) { *out_len = body + tail; return KEX_ERR_OUT_CAP; }
    for (uint32_t k = 0; k < h.npieces; ++k) {
      const Piece &pc = ph.pieces[h.piece_off + k];
      k_copy_tail<<<1, 64, 0, st>>>(P.consts + pc.off, pc.len, d_out + o);
      p->launches++;
      o += pc.len;
    }
    *out_len = o;
    *status = KEX_ACCEPT;
    *fail_count = 0;
  } else {
    // whole 16 KiB flushes only (crt.c:140-159, 217-227) -- unless this is the transducer phase of a stage
    // with register actions: its stream is not stdout but the interpreter's input, everything up to the
    // failing symbol is interpreted and the flush rule applies once, to the stage's output (kex_run_device)
    *out_len = feeds_interpreter ? body : body / 16384 * 16384;
    *status = KEX_REJECT;
    *fail_count = n_eff;
  }
  return KEX_OK;
}

extern "C" int kex_run_device(kex_program *p, const uint8_t *d_in, size_t n, uint8_t *d_out, size_t out_cap,
                              size_t *out_len, int *status, size_t *fail_count, void *stream) {
  if (!p || !out_len || !status || !fail_count || (n && !d_in)) return KEX_ERR_ARG;
  CK(cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;
  p->launches = 0;
  for (int i = 0; i < 4; ++i) p->ms[i] = 0.f;
  if (p->timing) CK(cudaEventRecord(p->ev[6], st));
  const uint8_t *cur = d_in;
  size_t cur_n = n;
  const size_t first = p->only_phase ? p->only_phase - 1 : 0;
  const size_t nph = p->only_phase ? p->only_phase : p->phases.size();
  for (size_t i = first; i < nph; ++i) {
    uint8_t *dst = d_out;
    size_t cap = out_cap;
    if (i + 1 < nph) {
      // intermediate stream: grow until it fits (the exact size is known before emit)
      Buf &b = p->inter[i & 1];
      size_t guess = cur_n * 2 + 65536;
      if (b.cap < guess) { int rc = ensure(p, b, guess); if (rc) return rc; }
      dst = (uint8_t *)b.p;
      cap = b.cap;
    }
    size_t ol = 0, fc = 0;
    int stt = 0;
    const bool feeds = i + 1 < nph && p->phases[i + 1].act;
    int rc = run_phase(p, (uint32_t)i, cur, cur_n, dst, cap, &ol, &stt, &fc, st, feeds);
    if (rc == KEX_ERR_OUT_CAP && i + 1 < nph) {
      Buf &b = p->inter[i & 1];
      CK(cudaStreamSynchronize(st));
      if ((rc = ensure(p, b, ol + 65536))) return rc;
      dst = (uint8_t *)b.p;
      cap = b.cap;
      rc = run_phase(p, (uint32_t)i, cur, cur_n, dst, cap, &ol, &stt, &fc, st, feeds);
    }
    if (rc) { *out_len = ol; return rc; }
    if (p->phases[i].act && i > first && *status == KEX_REJECT) {
      // the two phases of a stage with register actions are one stage of the
      // program: a reject of its transducer phase is the stage's reject, and
      // only whole 16 KiB flushes of the interpreted stream leave it
      stt = KEX_REJECT;
      fc = *fail_count;
      ol = ol / 16384 * 16384;
    }
    // a rejecting phase hands its truncated stream to the next phase, and the
    // pipeline's exit status is the last phase's (crt.c:414-455)
    *status = stt;
    *fail_count = fc;
    *out_len = ol;
    cur = dst;
    cur_n = ol;
  }
  if (p->timing) {
    CK(cudaEventRecord(p->ev[7], st));
    CK(cudaStreamSynchronize(st));
    if (p->c->sh_n) {
      cudaEventElapsedTime(&p->ms[0], p->ev[0], p->ev[1]);
      cudaEventElapsedTime(&p->ms[1], p->ev[2], p->ev[3]);
      if (*out_len) cudaEventElapsedTime(&p->ms[2], p->ev[4], p->ev[5]);
    }
    cudaEventElapsedTime(&p->ms[3], p->ev[6], p->ev[7]);
  }
  CK(cudaStreamSynchronize(st));
  return KEX_OK;
}

// ------------------------------------------------------------ block streaming
// stdin -> stdout with bounded memory (the role of the 2 x 16 KiB input window
// and the 16 KiB output window of crt/crt.c:61,299-305,107-136): the caller
// feeds blocks of the input in order; every block is evaluated as the next
// shard of one run (state map -> true start state -> seam summary), and the
// block before it is emitted once the new block's seam summary fixes the seam
// code at its end.  Single-phase programs on the v3 kernels whose seam
// summaries are constant maps (record-shaped programs).
// Emit block `blk` (its scratch context is cx[blk & 1]) with the seam code at its end.
static int stream_emit_block(kex_program *p, uint32_t blk, size_t n_eff, uint32_t code, uint8_t *h_out, size_t out_cap,
                             size_t *out_len) {
  p->c = &p->cx[blk & 1];
  *out_len = 0;
  // device buffer: a guess from the input size first, the exact size if that was too small
  int rc = ensure(p, p->sout, 4 * n_eff + (1u << 20));
  if (rc) return rc;
  size_t ol = 0;
  rc = do_emit(p, code, n_eff, (uint8_t *)p->sout.p, p->sout.cap - 16, &ol, nullptr);
  if (rc == KEX_ERR_OUT_CAP) {
    if ((rc = ensure(p, p->sout, ol + 16))) return rc;
    rc = do_emit(p, code, n_eff, (uint8_t *)p->sout.p, p->sout.cap - 16, &ol, nullptr);
  }
  if (rc) { *out_len = ol; return rc; }
  *out_len = ol;
  if (ol > out_cap) return KEX_ERR_OUT_CAP;
  if (ol) CK(cudaMemcpy(h_out, p->sout.p, ol, cudaMemcpyDeviceToHost));
  return KEX_OK;
}

extern "C" int kex_stream_begin(kex_program *p) {
  if (!p) return KEX_ERR_ARG;
  if (p->phases.size() != 1 || !p->phases[0].v3.ok) return KEX_ERR_UNSUPPORTED;
  CK(cudaSetDevice(p->device));
  p->st_open = true; p->st_failed = false;
  p->st_state = p->phases[0].dev.init; p->st_blocks = 0; p->st_nq = 0;
  p->st_consumed = 0; p->st_fail_at = 0;
  return KEX_OK;
}

extern "C" int kex_stream_feed(kex_program *p, const uint8_t *h_in, size_t n, uint8_t *h_out, size_t out_cap,
                               size_t *out_len) {
  if (!p || !p->st_open || !out_len || !h_in || n == 0) return KEX_ERR_ARG;
  CK(cudaSetDevice(p->device));
  *out_len = 0;
  if (p->st_failed) return KEX_OK;                       // input after the failure is ignored (C.hs:79-81)
  if (p->st_nq == 2) return KEX_ERR_UNSUPPORTED;         // two blocks already wait for their seam codes
  PhaseHost &ph = p->phases[0];
  const uint32_t NL = ph.fdev.NL;
  const uint32_t k = p->st_blocks;
  Buf &in = p->sin[k & 1];
  int rc = ensure(p, in, n + 16);
  if (rc) return rc;
  CK(cudaMemcpy(in.p, h_in, n, cudaMemcpyHostToDevice));
  p->c = &p->cx[k & 1];
  p->launches = 0;
  if ((rc = do_summarize(p, 0, (const uint8_t *)in.p, n, nullptr))) return rc;
  uint32_t end_state = 0;
  size_t fpos = (size_t)-1;
  std::vector<uint8_t> seam(NL);
  if ((rc = shard_walk_seam(p, p->st_state, &end_state, &fpos, seam.data(), nullptr))) return rc;
  const bool failed = fpos != (size_t)-1;
  bool constant = true;
  for (uint32_t l = 1; l < NL; ++l) constant = constant && seam[l] == seam[0];
  if (p->st_nq == 1 && (constant || failed)) {
    // this block fixes the seam code at the end of the block before it (after a failure nothing is live)
    rc = stream_emit_block(p, k - 1, p->st_q[0], seam[0], h_out, out_cap, out_len);
    p->c = &p->cx[0];
    if (rc) return rc;                                   // nothing was committed: the call can be repeated
    p->st_nq = 0;
  } else if (p->st_nq == 1) {
    p->st_seam = seam;                                   // resolved by kex_stream_end (e.g. a short last block)
  }
  p->st_q[p->st_nq++] = failed ? fpos : n;
  p->st_blocks = k + 1;
  if (failed) { p->st_failed = true; p->st_fail_at = p->st_consumed + fpos; }
  p->st_consumed += n;
  p->st_state = end_state;
  p->c = &p->cx[0];
  return KEX_OK;
}

extern "C" int kex_stream_end(kex_program *p, uint8_t *h_out, size_t out_cap, size_t *out_len, int *status,
                              size_t *fail_count) {
  if (!p || !p->st_open || !out_len || !status || !fail_count) return KEX_ERR_ARG;
  CK(cudaSetDevice(p->device));
  PhaseHost &ph = p->phases[0];
  *out_len = 0;
  const int32_t fa = p->st_failed ? -1 : ph.fin[p->st_state];
  const bool accept = fa >= 0;
  const uint32_t last_code = accept ? final_code(ph, p->st_state) : 0u;
  size_t body = 0;
  for (uint32_t i = 0; i < p->st_nq; ++i) {
    const uint32_t blk = p->st_blocks - p->st_nq + i;
    const uint32_t code = (i + 1 == p->st_nq) ? last_code : (uint32_t)p->st_seam[last_code];
    size_t ol = 0;
    int rc = stream_emit_block(p, blk, p->st_q[i], code, h_out + body, out_cap > body ? out_cap - body : 0, &ol);
    p->c = &p->cx[0];
    if (rc) { *out_len = body + ol + (rc == KEX_ERR_OUT_CAP ? (size_t)1 << 20 : 0); return rc; }
    body += ol;
  }
  if (accept) {
    const ActHdr &h = ph.acts[fa];
    size_t o = body;
    for (uint32_t k = 0; k < h.npieces; ++k) {
      const Piece &pc = ph.pieces[h.piece_off + k];
      if (o + pc.len > out_cap) { *out_len = o + ((size_t)1 << 20); return KEX_ERR_OUT_CAP; }
      memcpy(h_out + o, ph.consts.data() + pc.off, pc.len);
      o += pc.len;
    }
    *out_len = o;
    *status = KEX_ACCEPT;
    *fail_count = 0;
  } else {
    // the caller keeps only whole 16 KiB flushes of the whole stream (crt.c:140-159,217-227)
    *out_len = body;
    *status = KEX_REJECT;
    *fail_count = p->st_failed ? p->st_fail_at : p->st_consumed;
  }
  p->st_open = false;
  return KEX_OK;
}

// Host pipeline for single-phase programs on the v3 kernels: the input is cut
// into sub-waves that are copied in, evaluated as consecutive shards of one run
// (state map -> true start state -> seam summary -> emit, exactly the sharded
// entry points above) and copied out on three streams, so that the host<->device
// copies of neighbouring sub-waves overlap the kernels and each other.  A
// sub-wave is emitted as soon as the seam summary of its successor fixes the
// seam code at its end; if a summary is not a constant map the pipeline gives up
// and the whole input is evaluated at once.  Returns 1 when it gave up.
static int run_host_pipelined(kex_program *p, const uint8_t *h_in, size_t n, uint8_t *h_out, size_t out_cap,
                              size_t *out_len, int *status, size_t *fail_count, size_t wave) {
  PhaseHost &ph = p->phases[0];
  const uint32_t NL = ph.fdev.NL;
  if (!p->s_comp) {
    CK(cudaStreamCreateWithFlags(&p->s_h2d, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&p->s_comp, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&p->s_d2h, cudaStreamNonBlocking));
  }
  // sub-wave boundaries: a quarter and a half wave first and last, so that the first kernel starts
  // after a short copy and the last copy out is short (pipeline fill and drain)
  std::vector<size_t> cut;
  {
    const size_t q = (wave / 4) & ~(size_t)4095, h = (wave / 2) & ~(size_t)4095;
    size_t pos = 0;
    cut.push_back(0);
    const bool taper = q >= (1u << 20) && n >= 6 * wave;
    if (taper) { cut.push_back(pos += q); cut.push_back(pos += h); }
    const size_t tail = taper ? q + h : 0;
    while (n - pos > wave + tail) cut.push_back(pos += wave);
    if (taper) {
      const size_t mid = (n - q - h) & ~(size_t)4095;      // sub-waves start 4096-aligned
      if (mid > pos) cut.push_back(pos = mid);
      cut.push_back(pos += h);
    }
    cut.push_back(n);
  }
  const size_t nw = cut.size() - 1;
  while (p->pipe_ev.size() < 2 * nw) {
    cudaEvent_t e;
    CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    p->pipe_ev.push_back(e);
  }
  uint8_t *d_in = (uint8_t *)p->hostio_in.p, *d_out = (uint8_t *)p->hostio_out.p;
  for (size_t i = 0; i < nw; ++i) {
    const size_t off = cut[i], len = cut[i + 1] - cut[i];
    CK(cudaMemcpyAsync(d_in + off, h_in + off, len, cudaMemcpyHostToDevice, p->s_h2d));
    CK(cudaEventRecord(p->pipe_ev[2 * i], p->s_h2d));
  }
  cudaStream_t st = p->s_comp;
  std::vector<uint8_t> seam(NL);
  uint32_t state = ph.dev.init;
  size_t out_off = 0, n_fail = (size_t)-1;
  long pending = -1;                    // sub-wave walked but not yet emitted
  size_t pending_neff = 0;
  int rc = KEX_OK;
  bool gave_up = false, failed = false;
  uint32_t launches = 0;
  auto emit_one = [&](size_t k, uint32_t code, size_t n_eff) -> int {
    p->c = &p->cx[k & 1];
    p->emit_out_off = out_off;
    size_t ol = 0;
    p->launches = 0;
    int r = do_emit(p, code, n_eff, d_out, out_cap, &ol, st);
    launches += p->launches;
    p->emit_out_off = 0;
    if (r) return r;
    if (ol) {
      if (cudaEventRecord(p->pipe_ev[2 * k + 1], st) != cudaSuccess) return KEX_ERR_CUDA;
      if (cudaStreamWaitEvent(p->s_d2h, p->pipe_ev[2 * k + 1], 0) != cudaSuccess) return KEX_ERR_CUDA;
      if (cudaMemcpyAsync(h_out + out_off, d_out + out_off, ol, cudaMemcpyDeviceToHost, p->s_d2h) != cudaSuccess)
        return KEX_ERR_CUDA;
    }
    out_off += ol;
    return KEX_OK;
  };
  for (size_t i = 0; i < nw && rc == KEX_OK && !gave_up && !failed; ++i) {
    const size_t off = cut[i], len = cut[i + 1] - cut[i];
    p->c = &p->cx[i & 1];
    CK(cudaStreamWaitEvent(st, p->pipe_ev[2 * i], 0));
    p->launches = 0;
    if ((rc = do_summarize(p, 0, d_in + off, len, st))) break;
    uint32_t end_state = 0;
    size_t fpos = (size_t)-1;
    if ((rc = shard_walk_seam(p, state, &end_state, &fpos, seam.data(), st))) break;
    launches += p->launches;
    size_t n_eff = len;
    if (fpos != (size_t)-1) { failed = true; n_eff = fpos; n_fail = off + fpos; }
    if (pending >= 0) {
      // seam code at the end of the pending sub-wave: this sub-wave's summary must be a constant map,
      // or the run ends here (failure: nothing is live any more)
      bool constant = true;
      for (uint32_t l = 1; l < NL; ++l) constant = constant && seam[l] == seam[0];
      if (!constant && !failed) { gave_up = true; break; }
      if ((rc = emit_one((size_t)pending, seam[0], pending_neff))) break;    // (after a failure: the code of "nothing live")
      pending = -1;
    }
    pending = (long)i;
    pending_neff = n_eff;
    state = end_state;
  }
  if (rc == KEX_OK && !gave_up && pending >= 0) {
    // the last sub-wave: the end-of-input action (or the failure) fixes its seam code
    p->c = &p->cx[pending & 1];
    const int32_t fa = failed ? -1 : ph.fin[state];
    const bool accept = fa >= 0;
    const uint32_t code = accept ? final_code(ph, state) : 0u;
    rc = emit_one((size_t)pending, code, pending_neff);
    if (rc == KEX_OK) {
      if (accept) {
        const ActHdr &h = ph.acts[fa];
        size_t tail = 0;
        for (uint32_t k = 0; k < h.npieces; ++k) tail += ph.pieces[h.piece_off + k].len;
        if (out_off + tail > out_cap) {
          rc = KEX_ERR_OUT_CAP;
          *out_len = out_off + tail;
        } else {
          size_t o = out_off;
          for (uint32_t k = 0; k < h.npieces; ++k) {
            const Piece &pc = ph.pieces[h.piece_off + k];
            k_copy_tail<<<1, 64, 0, st>>>(ph.dev.consts + pc.off, pc.len, d_out + o);
            launches++;
            o += pc.len;
          }
          if (tail) {
            CK(cudaStreamSynchronize(st));
            CK(cudaMemcpyAsync(h_out + out_off, d_out + out_off, tail, cudaMemcpyDeviceToHost, p->s_d2h));
          }
          *out_len = o;
          *status = KEX_ACCEPT;
          *fail_count = 0;
        }
      } else {
        *out_len = out_off / 16384 * 16384;   // whole 16 KiB flushes only (crt.c:140-159, 217-227)
        *status = KEX_REJECT;
        *fail_count = failed ? n_fail : n;
      }
    }
  }
  p->c = &p->cx[0];
  cudaStreamSynchronize(p->s_h2d);
  cudaStreamSynchronize(st);
  cudaStreamSynchronize(p->s_d2h);
  p->launches = launches;
  if (gave_up) return 1;
  return rc;
}

extern "C" int kex_run_host(kex_program *p, const uint8_t *h_in, size_t n, uint8_t *h_out, size_t out_cap,
                            size_t *out_len, int *status, size_t *fail_count) {
  if (!p || !out_len || !status || !fail_count || (n && !h_in)) return KEX_ERR_ARG;
  CK(cudaSetDevice(p->device));
  int rc;
  if ((rc = ensure(p, p->hostio_in, n + 16))) return rc;
  if ((rc = ensure(p, p->hostio_out, out_cap + 16))) return rc;
  size_t wave = 64u << 20;
  if (const char *e = getenv("KEX_HOST_WAVE_MIB")) { const long x = atol(e); if (x > 0) wave = (size_t)x << 20; }
  if (p->phases.size() == 1 && p->phases[0].v3.ok && n >= 2 * wave && !getenv("KEX_NO_HOST_PIPELINE")) {   // (a selected phase of a 1-phase program is that phase)
    rc = run_host_pipelined(p, h_in, n, h_out, out_cap, out_len, status, fail_count, wave);
    if (rc <= 0 && rc != KEX_ERR_OUT_CAP) return rc;
    // the pipeline gave up (or the output did not fit): evaluate the resident input at once
    rc = kex_run_device(p, (const uint8_t *)p->hostio_in.p, n, (uint8_t *)p->hostio_out.p, out_cap, out_len, status,
                        fail_count, nullptr);
    if (rc) return rc;
    if (*out_len) CK(cudaMemcpy(h_out, p->hostio_out.p, *out_len, cudaMemcpyDeviceToHost));
    return KEX_OK;
  }
  if (n) CK(cudaMemcpy(p->hostio_in.p, h_in, n, cudaMemcpyHostToDevice));
  rc = kex_run_device(p, (const uint8_t *)p->hostio_in.p, n, (uint8_t *)p->hostio_out.p, out_cap, out_len, status,
                      fail_count, nullptr);
  if (rc) return rc;
  if (*out_len) CK(cudaMemcpy(h_out, p->hostio_out.p, *out_len, cudaMemcpyDeviceToHost));
  return KEX_OK;
}
