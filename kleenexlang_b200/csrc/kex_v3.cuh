// kex_v3.cuh -- warp-autonomous monoid kernels ("v3").
//
// Same algorithm as kex_fast.cuh (forward transition monoid, live-set monoid,
// the tile that creates a byte writes it), re-laid for the B200 SM:
//
//   k3_fwd     state chain of `matchN` (src/KMC/Program/Backends/C.hs:72-83)
//              for all start states at once.  One warp walks a pair of 4 KiB
//              chunks, lane = one 128-byte block of each (two independent
//              dependent-load chains, every warp load inside one contiguous
//              4 KiB region), two input bytes per table lookup (pair table: row
//              offset of the product element), prefix element sampled every 16
//              bytes relative to the block; the blocks' elements are composed
//              across the warp through the element composition table.
//   k3_seams   per 1 KiB tile from its true start state: backward (live-set)
//              element until it is a constant map; exact position of a failing
//              transition (C.hs:79-81) by atomicMin.
//   k3_emit    outputconst/outputarray/output and the register buffers of
//              crt/crt.c:161-283.  One persistent CTA per SM = 31 worker warps
//              + 1 scan warp, no CTA-wide barrier in the loop.  A worker owns a
//              1 KiB tile; each lane owns 32 input bytes as two independent
//              16-byte halves (ILP 2 in every pass).  Transition and emission
//              tables are replicated once per lane in shared memory (entry
//              stride 128 B: lane l only ever touches bank l, so the per-byte
//              lookups are bank-conflict free) and hold absolute shared
//              addresses.  Passes per byte: forward (action id [, backward
//              element]), count (IDP.4A accumulates length and record count at
//              once), write (input bytes into a swizzled per-warp staging
//              window at tile-relative offsets, one record per template), then
//              templates are copied word-wise.  The live set at the end of
//              every half is guessed from the state there and verified by the
//              count pass (exact re-evaluation on a miss).  The scan warp
//              chains one descriptor per 31-tile group with a decoupled
//              look-back and hands every worker its global offset, which the
//              worker only needs for the stage-out -- deferred until after the
//              next tile's count pass -- 16-byte stores aligned to the
//              destination, source words funnel-shifted from the window.
#pragma once

#define V3_TILE 1024u
#define V3_SUB 16u
#define V3_SPC (KEX_CHUNK / V3_SUB)     // samples per 4 KiB chunk (256)
#define V3_SPT (V3_TILE / V3_SUB)       // samples per tile (64)
#define V3_TPC (KEX_CHUNK / V3_TILE)    // tiles per chunk (4)
#define V3_RECCAP 192u                  // template records per tile kept in shared memory (first run; then adaptive)
#define V3_SMEM_MAX (227 * 1024)        // dynamic shared memory per CTA on sm_100

struct V3Dev {
  uint32_t ok;                 // phase runs on the v3 kernels
  uint32_t log;                // log2 of the table entry stride in bytes: 7 = one copy per lane, 5 = one per 8 lanes
  uint32_t o_mulB, o_trans, o_BE, o_cls, o_compB, o_applyB, o_tpl2, o_pool, o_slots, o_guess, o_warp;   // byte offsets in dynamic smem
  uint32_t pool_stride;
  uint32_t NE;                 // NL * A emission entries
  const uint32_t *be3;         // [NE]  len | T << 15 | S << 16 | (lam_before * A) << 24;  S: emits exactly the input byte
  const uint32_t *be3w;        // [NE]  write-pass copy (has_lit): a one-byte ASCII literal sits in bits 8-14 instead of a template
  uint32_t has_lit;
  uint32_t rmw_words;          // > 0: every template record is >= 3 bytes long: edge words merged (v3_template_rmw), at most this many words per template
  const uint32_t *tpl2;        // [NE]  pool offset | length << 16 | (hole offset + 1) << 24   (entries with T)
  // forward pass
  uint32_t pair;               // 1: fwdtab is the pair table [NM][C*C], 0: [NM][C]
  uint32_t rowbytes;           // bytes per row of fwdtab
  uint32_t recip;              // row offset -> element id: __umulhi(off, recip)
  uint32_t fwd_entries;
  const uint16_t *fwdtab;      // row byte offset of the product element
  const uint16_t *compF;       // [NM][NM] element id of (first, then second)
  uint32_t NM;
};

__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32_v(uint32_t a) {       // data written by this kernel
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds_u8_v(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
// inclusive warp scan step: the shuffle's own predicate says whether the source lane exists (SHFL + one
// predicated add instead of SHFL + compare + select + add)
__device__ __forceinline__ uint32_t v3_scan_step(uint32_t x, int d) {
  uint32_t r;
  asm volatile("{ .reg .u32 t; .reg .pred p; shfl.sync.up.b32 t|p, %1, %2, 0, 0xffffffff; @p add.u32 t, t, %1; mov.u32 %0, t; }"
               : "=r"(r) : "r"(x), "r"(d));
  return r;
}
// The live-set guess table of k3_emit is read and written without
// synchronisation on purpose (any value is only a guess that the count pass
// verifies); built with -DKEX_SANITIZE the helpers are not inlined, which keeps
// these accesses apart in compute-sanitizer reports (profiles/r01_sanitizer.txt).
#ifdef KEX_SANITIZE
#define KEX_GUESS_FN __noinline__
#else
#define KEX_GUESS_FN __forceinline__
#endif
__device__ KEX_GUESS_FN uint32_t guess_ld(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ KEX_GUESS_FN void guess_st(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
// acc - (byte 0 of e): unsigned bytes of e times signed bytes (-1, 0, 0, 0)
__device__ __forceinline__ uint32_t sub_byte0(uint32_t e, uint32_t acc) {
  int r;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(e), "r"(0x000000FF), "r"((int)acc));
  return (uint32_t)r;
}
__device__ __forceinline__ void ld_stream32(const uint8_t *p, uint32_t (&r)[8]) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}

// ------------------------------------------------------------------ k3_fwd
// Shared memory: clsA[256] (u8, first byte of a pair: class * C * 2; single
// steps: class * 2), clsB[256] (second byte: class * 2), then the table.
#define F3_PAIR(m, w, k)                                                                          \
  m = *(const uint16_t *)(tab + m + clsA[byte_at(w, k)] + clsB[byte_at(w, (k) + 1)]);
#define F3_ONE(m, w, k) m = *(const uint16_t *)(tab + m + clsA[byte_at(w, k)]);
#define F3_WORD(m, w)                                                                             \
  if (PAIR) { F3_PAIR(m, w, 0) F3_PAIR(m, w, 2) } else { F3_ONE(m, w, 0) F3_ONE(m, w, 1) F3_ONE(m, w, 2) F3_ONE(m, w, 3) }
#define F3_16(m, v) F3_WORD(m, v.x) F3_WORD(m, v.y) F3_WORD(m, v.z) F3_WORD(m, v.w)

// One warp per pair of 4 KiB chunks: lane l walks block l (128 bytes) of both
// chunks (two independent chains), so every load instruction of the warp reads
// inside one contiguous 4 KiB region.  Within a block the prefix element before
// every 16 bytes is sampled relative to the block start; the blocks' elements
// are composed across the warp (element composition table in L2) into the
// prefix element before every block (`blockpre`) and the chunk's state map.
//   state before sub-chunk j of chunk c = applyF[samples[c][j]][ applyF[blockpre[c][j / 8]][chunk start] ]
template <bool PAIR>
__device__ __forceinline__ uint32_t f3_block_tail(const PhaseDev &P, const FastDev &F, const V3Dev &V,
                                                  const uint8_t *clsA, const uint8_t *clsB, const uint8_t *tab,
                                                  const uint8_t *__restrict__ p, uint32_t len,
                                                  uint16_t *__restrict__ srow) {
  uint32_t m = 0;
  const uint32_t C = P.C;
  for (uint32_t j = 0; j < len;) {
    if ((j & (V3_SUB - 1u)) == 0) srow[j / V3_SUB] = (uint16_t)__umulhi(m, V.recip);
    if (PAIR) {
      if (j + 1 < len) {
        m = *(const uint16_t *)(tab + m + clsA[p[j]] + clsB[p[j + 1]]);
        j += 2;
      } else {
        const uint32_t id = __umulhi(m, V.recip);
        m = (uint32_t)F.mulF[id * C + P.cls[p[j]]] * V.rowbytes;
        j += 1;
      }
    } else {
      m = *(const uint16_t *)(tab + m + clsA[p[j]]);
      j += 1;
    }
  }
  return __umulhi(m, V.recip);
}

// inclusive composition over the warp: lane l gets e_0 . e_1 ... e_l (earlier first)
__device__ __forceinline__ uint32_t f3_warp_compose(const uint16_t *__restrict__ compF, uint32_t NM, uint32_t e,
                                                    uint32_t lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, e, d);
    if (lane >= (uint32_t)d) e = __ldg(compF + (size_t)y * NM + e);
  }
  return e;
}

#define V3_BLK 128u                      // bytes per lane and chunk
#define V3_BPC (KEX_CHUNK / V3_BLK)      // blocks per chunk (32 = one warp)

template <bool PAIR>
__global__ void __launch_bounds__(512, 2)
k3_fwd(PhaseDev P, FastDev F, V3Dev V, const uint8_t *__restrict__ in, size_t n, size_t nchunks,
       uint16_t *__restrict__ samples, uint16_t *__restrict__ blockpre, uint16_t *__restrict__ maps) {
  extern __shared__ __align__(16) uint8_t smem_f3[];
  uint8_t *clsA = smem_f3, *clsB = smem_f3 + 256;
  uint8_t *tab = smem_f3 + 512;
  const uint32_t C = P.C, Q1 = P.Q + 1, NM = V.NM;
  for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
    const uint32_t c = P.cls[i];
    clsA[i] = (uint8_t)(PAIR ? c * C * 2u : c * 2u);
    clsB[i] = (uint8_t)(c * 2u);
  }
  for (uint32_t i = threadIdx.x; i < V.fwd_entries; i += blockDim.x) ((uint16_t *)tab)[i] = V.fwdtab[i];
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u, wpc = blockDim.x >> 5;
  const size_t npairs = (nchunks + 1) / 2;
  const uint32_t recip = V.recip;
  const size_t gw = (size_t)blockIdx.x * wpc + (threadIdx.x >> 5), nw = (size_t)gridDim.x * wpc;
  for (size_t pr = gw; pr < npairs; pr += nw) {
    const size_t c0 = 2 * pr, c1 = c0 + 1;
    const size_t base0 = c0 * KEX_CHUNK;
    const bool two = c1 < nchunks;
    const uint32_t len0 = (uint32_t)((n - base0 < KEX_CHUNK) ? (n - base0) : KEX_CHUNK);
    const uint32_t len1 = two ? (uint32_t)((n - base0 - KEX_CHUNK < KEX_CHUNK) ? (n - base0 - KEX_CHUNK) : KEX_CHUNK) : 0u;
    uint16_t *sa = samples + c0 * V3_SPC + lane * (V3_BLK / V3_SUB), *sb = sa + V3_SPC;
    uint32_t ea, eb;                         // element ids of this lane's two blocks
    if (len0 == KEX_CHUNK && len1 == KEX_CHUNK) {
      const uint8_t *pa = in + base0 + lane * V3_BLK, *pb = pa + KEX_CHUNK;
      uint32_t ma = 0, mb = 0;
      uint32_t ca[8], cb[8], na[8], nb[8];
      ld_stream32(pa, ca);
      ld_stream32(pb, cb);
      uint32_t qa[4], qb[4];
#pragma unroll
      for (uint32_t blk = 0; blk < V3_BLK / 32u; ++blk) {
        const uint32_t nx = (blk + 1u < V3_BLK / 32u) ? (blk + 1u) * 32u : blk * 32u;
        ld_stream32(pa + nx, na);
        ld_stream32(pb + nx, nb);
        const uint32_t ia0 = ma, ib0 = mb;
        F3_WORD(ma, ca[0]) F3_WORD(mb, cb[0]) F3_WORD(ma, ca[1]) F3_WORD(mb, cb[1])
        F3_WORD(ma, ca[2]) F3_WORD(mb, cb[2]) F3_WORD(ma, ca[3]) F3_WORD(mb, cb[3])
        const uint32_t ia1 = ma, ib1 = mb;
        F3_WORD(ma, ca[4]) F3_WORD(mb, cb[4]) F3_WORD(ma, ca[5]) F3_WORD(mb, cb[5])
        F3_WORD(ma, ca[6]) F3_WORD(mb, cb[6]) F3_WORD(ma, ca[7]) F3_WORD(mb, cb[7])
        qa[blk] = __umulhi(ia0, recip) | (__umulhi(ia1, recip) << 16);
        qb[blk] = __umulhi(ib0, recip) | (__umulhi(ib1, recip) << 16);
#pragma unroll
        for (int k = 0; k < 8; ++k) { ca[k] = na[k]; cb[k] = nb[k]; }
      }
      *(uint4 *)sa = make_uint4(qa[0], qa[1], qa[2], qa[3]);
      *(uint4 *)sb = make_uint4(qb[0], qb[1], qb[2], qb[3]);
      ea = __umulhi(ma, recip);
      eb = __umulhi(mb, recip);
    } else {
      const uint32_t lo = lane * V3_BLK;
      const uint32_t l0 = (lo < len0) ? ((len0 - lo < V3_BLK) ? (len0 - lo) : V3_BLK) : 0u;
      const uint32_t l1 = (lo < len1) ? ((len1 - lo < V3_BLK) ? (len1 - lo) : V3_BLK) : 0u;
      ea = f3_block_tail<PAIR>(P, F, V, clsA, clsB, tab, in + base0 + lo, l0, sa);
      eb = f3_block_tail<PAIR>(P, F, V, clsA, clsB, tab, in + base0 + KEX_CHUNK + lo, l1, sb);
    }
    // prefix element before every block, element of the whole chunk
    const uint32_t xa = f3_warp_compose(V.compF, NM, ea, lane), xb = f3_warp_compose(V.compF, NM, eb, lane);
    uint32_t pa2 = __shfl_up_sync(0xFFFFFFFFu, xa, 1), pb2 = __shfl_up_sync(0xFFFFFFFFu, xb, 1);
    if (lane == 0) { pa2 = 0; pb2 = 0; }
    blockpre[c0 * V3_BPC + lane] = (uint16_t)pa2;
    const uint32_t ta = __shfl_sync(0xFFFFFFFFu, xa, 31), tb = __shfl_sync(0xFFFFFFFFu, xb, 31);
    for (uint32_t q = lane; q < Q1; q += 32u) maps[c0 * Q1 + q] = F.applyF[(size_t)ta * Q1 + q];
    if (two) {
      blockpre[c1 * V3_BPC + lane] = (uint16_t)pb2;
      for (uint32_t q = lane; q < Q1; q += 32u) maps[c1 * Q1 + q] = F.applyF[(size_t)tb * Q1 + q];
    }
  }
}

// ------------------------------------------------------------------ k3_seams
// One thread per 1 KiB tile.  Shared memory: cls4[256] (4 * class), transition
// table u32 [(Q+1)*C] = byte offset of the next state's row | byte offset of the
// generator inside a backward-element row << 16 | FAIL << 31, backward-element
// table u16 [NB][NG] = byte offset of the product's row | constant-map << 15.
// A step is LDS.U8 + LDS + LDS.U16 and a handful of integer instructions.
#define S3_STEP(word, k)                                                                          \
  {                                                                                               \
    const uint32_t e = *(const uint32_t *)(tr + srow + cls4[byte_at(word, k)]);                   \
    if ((int)e < 0) { fail_at = g + (uint32_t)(pos); goto seam_done; }                            \
    srow = e & 0xFFFFu;                                                                           \
    const uint32_t m = *(const uint16_t *)(mb2 + mrow + ((e >> 16) & 0x7FFFu));                   \
    mrow = m & 0x7FFFu;                                                                           \
    if ((m & 0x8000u) && !failing) goto seam_done;                                                \
  }
__global__ void __launch_bounds__(256)
k3_seams(PhaseDev P, FastDev F, const uint8_t *__restrict__ in, size_t n, size_t ntiles,
         const uint16_t *__restrict__ blockpre, const uint16_t *__restrict__ chunk_start,
         const uint16_t *__restrict__ maps, uint8_t *__restrict__ bmaps, RunResult *__restrict__ res) {
  extern __shared__ __align__(16) uint8_t smem_s3[];
  const uint32_t Q = P.Q, Q1 = Q + 1, C = P.C, NL = F.NL, NG = F.NG, NB = F.NB;
  uint8_t *cls4 = smem_s3;
  uint8_t *tr = smem_s3 + 256;
  uint8_t *mb2 = tr + 4u * Q1 * C;
  for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) cls4[i] = (uint8_t)(4u * P.cls[i]);
  for (uint32_t i = threadIdx.x; i < Q1 * C; i += blockDim.x) {
    const uint32_t e = F.trans2[i], nx = e & 0xFFFFu;
    ((uint32_t *)tr)[i] = (nx * C * 4u) | ((e >> 24) * 2u) << 16 | (nx == Q ? 0x80000000u : 0u);
  }
  for (uint32_t i = threadIdx.x; i < NB * NG; i += blockDim.x) {
    const uint32_t nx = F.mulB[i];
    ((uint16_t *)mb2)[i] = (uint16_t)((nx * NG * 2u) | (F.constB[nx] ? 0x8000u : 0u));
  }
  __syncthreads();
  const size_t tile = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tile >= ntiles) return;
  const size_t chunk = tile / V3_TPC;
  const uint32_t cs = chunk_start[chunk];
  uint32_t s = Q, endst = Q;
  if (cs != Q) {
    // a tile starts on a block boundary: its sample there is the identity
    s = F.applyF[(size_t)blockpre[tile * (V3_TILE / V3_BLK)] * Q1 + cs];
    const bool last_in_chunk = ((tile + 1) % V3_TPC == 0) || (tile + 1 == ntiles);
    endst = last_in_chunk ? maps[chunk * Q1 + cs] : F.applyF[(size_t)blockpre[(tile + 1) * (V3_TILE / V3_BLK)] * Q1 + cs];
  }
  if (tile == ntiles - 1) res->end_state = endst;
  const bool failing = (s != Q) && (endst == Q);
  uint32_t mrow = 0, fail_at = KEX_NONE32;
  const size_t base = tile * V3_TILE;
  if (s != Q && (failing || NL > 1)) {
    const uint32_t len = (uint32_t)((n - base < V3_TILE) ? (n - base) : V3_TILE);
    const uint8_t *p = in + base;
    uint32_t srow = s * C * 4u;
    uint32_t g = 0;
    for (; g + 16u <= len; g += 16u) {
      const uint4 v = *(const uint4 *)(p + g);
#define pos 0
      S3_STEP(v.x, 0)
#undef pos
#define pos 1
      S3_STEP(v.x, 1)
#undef pos
#define pos 2
      S3_STEP(v.x, 2)
#undef pos
#define pos 3
      S3_STEP(v.x, 3)
#undef pos
#define pos 4
      S3_STEP(v.y, 0)
#undef pos
#define pos 5
      S3_STEP(v.y, 1)
#undef pos
#define pos 6
      S3_STEP(v.y, 2)
#undef pos
#define pos 7
      S3_STEP(v.y, 3)
#undef pos
#define pos 8
      S3_STEP(v.z, 0)
#undef pos
#define pos 9
      S3_STEP(v.z, 1)
#undef pos
#define pos 10
      S3_STEP(v.z, 2)
#undef pos
#define pos 11
      S3_STEP(v.z, 3)
#undef pos
#define pos 12
      S3_STEP(v.w, 0)
#undef pos
#define pos 13
      S3_STEP(v.w, 1)
#undef pos
#define pos 14
      S3_STEP(v.w, 2)
#undef pos
#define pos 15
      S3_STEP(v.w, 3)
#undef pos
    }
    for (; g < len; ++g) {
      const uint32_t w1 = p[g];
#define pos 0
      S3_STEP(w1, 0)
#undef pos
    }
  seam_done:;
  }
  if (fail_at != KEX_NONE32) atomicMin(&res->fail_pos, (unsigned long long)(base + fail_at));
  if (NL > 1) {
    const uint32_t mb = mrow / (NG * 2u);
    for (uint32_t l = 0; l < NL; ++l) bmaps[tile * NL + l] = __ldg(F.applyB + mb * NL + l);
  }
}

// ------------------------------------------------------------------ k3_emit
// Workers wait for the scan warp's offsets on a named barrier from three places (deferred stage-out,
// window exceeded, after the loop).  PTX only asks for warp-level convergence at a `bar.sync`, but
// compute-sanitizer's synccheck reports warps of one block meeting the same barrier at different
// program counters; through this out-of-line function they all meet it at one.
__device__ __noinline__ void ef_wait_bases(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

extern __shared__ __align__(1024) uint8_t smem_v3[];

// Staging windows are addressed through this swizzle: address bits 7-11 (the
// 128-byte row) are XORed into bits 2-6 (the bank), so that lanes whose outputs
// lie a multiple of 64 bytes apart -- the usual case, a lane's 32 input bytes
// become about 64 output bytes -- hit different banks.  Words stay intact.
__device__ __forceinline__ uint32_t swz(uint32_t a) { return a ^ ((a >> 5) & 0x7Cu); }

// put byte 0 of e into byte k of acc
__device__ __forceinline__ uint32_t put_byte0(uint32_t acc, uint32_t e, int k) {
  const uint32_t sel = k == 0 ? 0x3214u : k == 1 ? 0x3240u : k == 2 ? 0x3410u : 0x4210u;
  return __byte_perm(acc, e, sel);
}

// Forward step.  E: last transition entry (bits 16-31: absolute shared address
// of the current state's row in this lane's copy; byte 0: action; byte 1:
// 2 * backward generator).  MB: absolute address of the backward element's row
// (256-byte rows of u16 row addresses).
template <int LOG, bool REGS>
__device__ __forceinline__ void v3_fstep(uint32_t cls_abs, uint32_t &E, uint32_t &MB, uint32_t w, uint32_t &ap, int k) {
  // the class table is 256-byte aligned: one PRMT extracts the byte and adds the table address
  const uint32_t c = lds_u8(__byte_perm(w, cls_abs, 0x7650u | (uint32_t)k));
  E = lds_u32((E >> 16) + (c << LOG));
  ap = put_byte0(ap, E, k);
  if (REGS) MB = lds_u16(__byte_perm(MB, E, 0x3215u));
}

template <int LOG, bool REGS, bool FULL>
__device__ __forceinline__ void v3_forward(uint32_t cls_abs, const uint32_t (&w)[8], uint32_t cnt_pos, uint32_t &EA,
                                           uint32_t &EB, uint32_t &MA, uint32_t &MB, uint32_t (&ap)[8]) {
#pragma unroll
  for (int k = 0; k < 8; ++k) ap[k] = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    if (FULL || (uint32_t)j < cnt_pos) v3_fstep<LOG, REGS>(cls_abs, EA, MA, w[j >> 2], ap[j >> 2], j & 3);
    if (FULL || (uint32_t)(j + 16) < cnt_pos) v3_fstep<LOG, REGS>(cls_abs, EB, MB, w[4 + (j >> 2)], ap[4 + (j >> 2)], j & 3);
  }
}

// Emission entry address of action a under the live set encoded in EL (byte 3
// = lam * A; bits 24-LOG .. 23 are zero, so EL >> (24-LOG) = (lam * A) << LOG).
// (one IDP.4A on the fma pipe extracts action byte k, scales it by the entry
// stride and adds the table base; the alu pipe is the busier one in these passes)
template <int LOG>
__device__ __forceinline__ uint32_t v3_be_addr(uint32_t pbe, uint32_t EL, uint32_t apw, int k) {
  return __dp4a(apw, (1u << LOG) << (8 * k), pbe) + (EL >> (24 - LOG));
}

// count pass over both halves: byte length in the low 14 bits, records above
template <int LOG>
__device__ __forceinline__ void v3_count(uint32_t pbe, const uint32_t (&ap)[8], uint32_t &ELA, uint32_t &ELB,
                                         uint32_t &accA, uint32_t &accB) {
#pragma unroll
  for (int j = 15; j >= 0; --j) {
    ELA = lds_u32(v3_be_addr<LOG>(pbe, ELA, ap[j >> 2], j & 3));
    accA = __dp4a(ELA, 0x00008001u, accA);
    ELB = lds_u32(v3_be_addr<LOG>(pbe, ELB, ap[4 + (j >> 2)], j & 3));
    accB = __dp4a(ELB, 0x00008001u, accB);
  }
}

template <int LOG, bool LIT>
__device__ __forceinline__ void v3_wstep(uint32_t pbe, uint32_t &EL, uint32_t apw, uint32_t w, int k, uint32_t &o,
                                         uint32_t &recp) {
  const uint32_t addr = v3_be_addr<LOG>(pbe, EL, apw, k);
  EL = lds_u32(addr);
  o = sub_byte0(EL, o);
  const uint32_t b = k == 0 ? w : __umulhi(w, 1u << (32 - 8 * k));     // w >> 8k on the fma pipe; the store takes the low byte
  if (EL & 0x10000u) sts_u8(swz(o), b);
  if (LIT && (EL & 0x7F00u)) sts_u8(swz(o), (EL >> 8) & 0x7Fu);      // one-byte literal: no template record
  if (EL & 0x8000u) {
    // staging address (18 bits) | entry address / 4 << 18; the input byte, for a template with a hole
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(recp), "r"(o + (addr << 16)), "r"(b) : "memory");
    recp += 8u;
  }
}

template <int LOG, bool LIT>
__device__ __forceinline__ void v3_write(uint32_t pbe, const uint32_t (&ap)[8], const uint32_t (&w)[8], uint32_t ELA,
                                         uint32_t ELB, uint32_t oA, uint32_t oB, uint32_t recpA, uint32_t recpB) {
#pragma unroll
  for (int j = 15; j >= 0; --j) {
    v3_wstep<LOG, LIT>(pbe, ELA, ap[j >> 2], w[j >> 2], j & 3, oA, recpA);
    v3_wstep<LOG, LIT>(pbe, ELB, ap[4 + (j >> 2)], w[4 + (j >> 2)], j & 3, oB, recpB);
  }
}

// One template record: the literal bytes go to the staging window word-wise
// (the pool is stored four times, copy v shifted by v bytes, so the unaligned
// source of an aligned destination word is one aligned load).
__device__ __forceinline__ void v3_copy_template(uint32_t pool_abs, uint32_t pool_stride, uint32_t o, uint32_t t,
                                                 uint32_t byte) {
  const uint32_t src = t & 0xFFFFu, len = (t >> 16) & 0xFFu;
  // head bytes up to the first aligned destination word, whole words, tail bytes:
  // straight-line and predicated, so the lanes of a warp do not diverge
  uint32_t head = (4u - (o & 3u)) & 3u;
  if (head > len) head = len;
  const uint32_t nw = (len - head) >> 2, tail = (len - head) & 3u;
  const uint32_t ps = pool_abs + src;
#pragma unroll
  for (uint32_t k = 0; k < 3; ++k)
    if (k < head) sts_u8(swz(o + k), lds_u8(ps + k));
  const uint32_t x = src + head;                       // pool offset of the first whole word
  const uint32_t wsrc = pool_abs + (x & 3u) * pool_stride + (x & ~3u);
  const uint32_t wdst = o + head;
#pragma unroll
  for (uint32_t i = 0; i < 6; ++i)
    if (i < nw) sts_u32(swz(wdst + 4u * i), lds_u32(wsrc + 4u * i));
  for (uint32_t i = 6; i < nw; ++i) sts_u32(swz(wdst + 4u * i), lds_u32(wsrc + 4u * i));
  const uint32_t tb = head + 4u * nw;
#pragma unroll
  for (uint32_t k = 0; k < 3; ++k)
    if (k < tail) sts_u8(swz(o + tb + k), lds_u8(ps + tb + k));
  if (t >> 24) sts_u8(swz(o + (t >> 24) - 1u), byte);  // the hole takes the input byte
}

// The same with merged edge words instead of edge bytes (every template of the program is at least three
// bytes long, so records two apart never share a word; neighbours are handled in different rounds): the first
// and the last word take one LOP3 select each, the words in between are plain copies; the loop is bounded by
// the program's longest template.  (Pool offsets include four bytes of padding: the word in front of a template exists.)
__device__ __forceinline__ uint32_t v3_bitsel(uint32_t a, uint32_t b, uint32_t m) {       // (a & m) | (b & ~m)
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(r) : "r"(m), "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ void v3_template_rmw(uint32_t pool_abs, uint32_t pool_stride, uint32_t o, uint32_t t,
                                                uint32_t byte, uint32_t nwmax) {
  const uint32_t src = t & 0xFFFFu, len = (t >> 16) & 0xFFu, hole = t >> 24;
  const uint32_t oe = o + len;
  const uint32_t x = src - (o & 3u), v = x & 3u;
  const uint32_t ps = pool_abs + v * pool_stride + (x - v);
  const uint32_t w0 = o & ~3u, nw = ((oe + 3u) >> 2) - (o >> 2);
  const uint32_t mlo = 0xFFFFFFFFu << (8u * (o & 3u)), mhi = 0xFFFFFFFFu >> (8u * ((0u - oe) & 3u));
  {
    const uint32_t m = (nw == 1u) ? (mlo & mhi) : mlo;
    const uint32_t da = swz(w0);
    sts_u32(da, v3_bitsel(lds_u32(ps), lds_u32_v(da), m));
  }
  for (uint32_t i = 1; i + 1u < nwmax; ++i)
    if (i + 1u < nw) sts_u32(swz(w0 + 4u * i), lds_u32(ps + 4u * i));
  if (nw > 1u) {
    const uint32_t da = swz(w0 + 4u * (nw - 1u));
    sts_u32(da, v3_bitsel(lds_u32(ps + 4u * (nw - 1u)), lds_u32_v(da), mhi));
  }
  if (hole) sts_u8(swz(o + hole - 1u), byte);
}

// Staging window -> global with 16-byte stores aligned to the destination:
// destination chunk c holds output bytes [16c - a, 16c - a + 16), a = gbase mod 16.
__device__ __forceinline__ void v3_stage_out(uint32_t stage_abs, uint32_t total, unsigned long long gbase,
                                             uint8_t *__restrict__ out, uint32_t lane) {
  const uint32_t a = (uint32_t)(gbase & 15ull);
  uint8_t *gal = out + (gbase - a);
  const uint32_t end = a + total;                             // in destination-chunk coordinates
  const uint32_t c_lo = a ? 1u : 0u, c_hi = end >> 4;         // whole chunks [c_lo, c_hi)
  const uint32_t sh = ((0u - a) & 3u) * 8u;
  for (uint32_t c = c_lo + lane; c < c_hi; c += 32u) {
    // source bytes start at window offset 16c - a: five words, funnel-shifted
    const uint32_t sa = stage_abs + ((16u * c - a) & ~3u);
    uint32_t W[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) W[k] = lds_u32_v(swz(sa + 4u * k));
    *(uint4 *)(gal + 16u * c) = make_uint4(__funnelshift_r(W[0], W[1], sh), __funnelshift_r(W[1], W[2], sh),
                                           __funnelshift_r(W[2], W[3], sh), __funnelshift_r(W[3], W[4], sh));
  }
  // head (destination bytes a..15 of chunk 0) and tail (after the last whole chunk)
  // (fewer than 16 bytes each: one predicated step, no loop)
  if (a) {
    const uint32_t he = (end < 16u) ? end : 16u;
    if (a + lane < he) gal[a + lane] = (uint8_t)lds_u8_v(swz(stage_abs + lane));
  }
  if (c_hi >= c_lo) {
    const uint32_t b2 = (c_hi << 4) + lane;
    if (b2 < end && b2 >= a) gal[b2] = (uint8_t)lds_u8_v(swz(stage_abs + b2 - a));
  }
}

// programs whose emissions are at most three bytes long: no word path
__device__ __forceinline__ void v3_copy_short(uint32_t pool_abs, uint32_t o, uint32_t t, uint32_t byte) {
  const uint32_t ps = pool_abs + (t & 0xFFFFu), len = (t >> 16) & 0xFFu;
#pragma unroll
  for (uint32_t k = 0; k < 3; ++k)
    if (k < len) sts_u8(swz(o + k), lds_u8(ps + k));
  if (t >> 24) sts_u8(swz(o + (t >> 24) - 1u), byte);
}

template <int LOG, bool REGS, bool LIT>
__global__ void __launch_bounds__(1024, 1)
k3_emit(PhaseDev P, FastDev F, V3Dev V, const uint8_t *__restrict__ in, size_t n_eff, uint32_t ntiles,
        const uint16_t *__restrict__ samples, const uint16_t *__restrict__ blockpre,
        const uint16_t *__restrict__ chunk_start, const uint8_t *__restrict__ lam_end,
        unsigned long long *__restrict__ desc, FastCtl *__restrict__ ctl, uint8_t *__restrict__ out, size_t out_cap,
        unsigned long long out_off, uint32_t stage_bytes,
        uint32_t warp_bytes, uint32_t reccap, uint32_t nospec) {
  constexpr uint32_t STRIDE = 1u << LOG, REP = STRIDE / 4u;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const uint32_t Q = P.Q, Q1 = Q + 1, C = P.C, A = P.A, NL = F.NL, NB = F.NB, NG = F.NG;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem_v3);
  const uint32_t slot4 = (lane & (REP - 1u)) * 4u;

  // ---- tables (once per CTA)
  const uint32_t mulB_abs = base + V.o_mulB, trans_abs = base + V.o_trans, be_abs = base + V.o_BE;
  if ((mulB_abs & 255u) || ((base + V.o_cls) & 255u) || be_abs + (LIT ? 2u : 1u) * V.NE * STRIDE > 65536u) __trap();
  // A row of the backward-element table is 256 bytes and holds `mcopies` copies of
  // its NG entries, (1 << mshift) bytes apart; lane l uses copy l % mcopies (the
  // copy offset is baked into byte 1 of its transition entries), so lanes in
  // different rows rarely meet in a bank.
  uint32_t mshift = 1;
  while ((1u << mshift) < 2u * NG) ++mshift;
  uint32_t mcopies = 128u >> mshift;            // the copies of a row span the 32 banks once
  if (mcopies > REP) mcopies = REP;
  if (mcopies < 1u) mcopies = 1u;
  for (uint32_t i = tid; i < NB * NG * mcopies; i += blockDim.x) {
    const uint32_t k = i % mcopies, j = i / mcopies, r = j / NG, g = j - r * NG;
    *(uint16_t *)(smem_v3 + V.o_mulB + r * 256u + (k << mshift) + 2u * g) = (uint16_t)(mulB_abs + (uint32_t)F.mulB[j] * 256u);
  }
  for (uint32_t i = tid; i < Q1 * C * REP; i += blockDim.x) {
    const uint32_t ent = i / REP, s = i - ent * REP;
    const uint32_t e = F.trans2[ent];
    const uint32_t row = trans_abs + (e & 0xFFFFu) * C * STRIDE + s * 4u;
    const uint32_t gb = 2u * (e >> 24) + ((s % mcopies) << mshift);
    *(uint32_t *)(smem_v3 + V.o_trans + ent * STRIDE + s * 4u) = (row << 16) | ((e >> 16) & 0xFFu) | (gb << 8);
  }
  for (uint32_t i = tid; i < V.NE * REP; i += blockDim.x) {
    const uint32_t ent = i / REP, s = i - ent * REP;
    *(uint32_t *)(smem_v3 + V.o_BE + ent * STRIDE + s * 4u) = V.be3[ent];
    if (LIT) *(uint32_t *)(smem_v3 + V.o_BE + (V.NE + ent) * STRIDE + s * 4u) = V.be3w[ent];
  }
  for (uint32_t i = tid; i < 256; i += blockDim.x) smem_v3[V.o_cls + i] = P.cls[i];
  for (uint32_t i = tid; i < Q1; i += blockDim.x) smem_v3[V.o_guess + i] = 0xFFu;
  for (uint32_t i = tid; i < NB * NB; i += blockDim.x) smem_v3[V.o_compB + i] = F.compB[i];
  for (uint32_t i = tid; i < NB * NL; i += blockDim.x) smem_v3[V.o_applyB + i] = F.applyB[i];
  for (uint32_t i = tid; i < V.NE; i += blockDim.x) *(uint32_t *)(smem_v3 + V.o_tpl2 + 4u * i) = V.tpl2[i];
  for (uint32_t i = tid; i < 4u * V.pool_stride; i += blockDim.x) {
    // copy v, byte i holds padded-pool byte i + v; the padded pool starts with 4 zero bytes (tpl2 offsets
    // include them), so that the word in front of a template exists in every copy (v3_template_rmw)
    const uint32_t v = i / V.pool_stride, k = i - v * V.pool_stride + v;
    smem_v3[V.o_pool + i] = (k >= 4u && k - 4u < F.pool_len) ? F.pool[k - 4u] : (uint8_t)0;
  }
  __syncthreads();

  const uint32_t cls_abs = base + V.o_cls, pool_abs = base + V.o_pool, tpl2_abs = base + V.o_tpl2;
  const uint32_t pbe = be_abs + slot4;
  const uint32_t bew_abs = be_abs + (LIT ? V.NE * STRIDE : 0u), pbw = bew_abs + slot4;     // write-pass copy of the emission table
  const uint32_t wreg = V.o_warp + warp * warp_bytes;            // this warp's staging window, then its records
  const uint32_t stage_abs = base + wreg;
  const uint32_t recs_abs = stage_abs + stage_bytes + 128u;    // window + one row of slack, both multiples of 128 (swizzle)
  const uint8_t *compB = smem_v3 + V.o_compB, *applyB = smem_v3 + V.o_applyB;
  // A group is the nwork consecutive tiles the worker warps of one CTA process
  // at a time.  The last warp of the CTA is the scan warp: it collects the tile
  // totals of the group from shared memory, chains ONE descriptor per group with
  // a decoupled look-back, and hands every worker its global output offset.  The
  // workers do not wait for it until their staging window is complete, so the
  // look-back latency hides behind the write pass.
  //   named barrier 1: workers arrive (totals published), scan warp waits
  //   named barrier 2: scan warp arrives (offsets published), workers wait
  const uint32_t nwork = nwarp - 1u;
  volatile uint32_t *slots = (volatile uint32_t *)(smem_v3 + V.o_slots);                             // [2][32] tile totals
  volatile unsigned long long *bases = (volatile unsigned long long *)(smem_v3 + V.o_slots + 256u);  // [2][32] output offsets
  const uint32_t ngroups = (ntiles + nwork - 1u) / nwork;
  const uint32_t bar_n = blockDim.x;
  uint32_t par = 0;
  const uint32_t guess_abs = base + V.o_guess;        // [Q+1] live set usually seen when a half ends in this state
  uint32_t spec_ctr = nospec ? 0x80000000u : 0u;     // sign bit: evaluate every tile exactly

  if (warp == nwork) {
    // =============================================================== scan warp
    for (uint32_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x, par ^= 1u) {
      asm volatile("bar.sync %0, %1;" ::"r"(1u + par), "r"(bar_n) : "memory");
      const uint32_t tv = (lane < nwork) ? slots[par * 32u + lane] : 0u;
      uint32_t inc_s = tv;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc_s, d);
        if (lane >= (uint32_t)d) inc_s += y;
      }
      const uint32_t gsum = __shfl_sync(0xFFFFFFFFu, inc_s, 31);
      if (lane == 0) st_desc(desc + grp, EF_FLAG_AGG | (unsigned long long)gsum);
      // every lane polls 8 descriptors (a window of 256 groups: all groups in flight
      // are covered by one round trip to L2), nearest first
      unsigned long long gex = 0;
      long long idx = (long long)grp - 1;
      bool ok = true;
      while (idx >= 0 && ok) {
        unsigned long long d[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const long long j = idx - (long long)(lane + 32u * k);
          d[k] = (j >= 0) ? ld_desc(desc + j) : EF_FLAG_INC;
        }
        bool hit = false;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (!hit) {
            const long long j = idx - (long long)(lane + 32u * k);
            if ((d[k] >> 62) == 0) {
              // a predecessor that has not published yet: poll; give up only after 20 s of wall time
              // (a descheduled predecessor on a time-sliced GPU must not fail a correct run)
              const unsigned long long t0 = ef_globaltimer();
              uint32_t spins = 0;
              while ((d[k] >> 62) == 0) {
                if ((++spins & 1023u) == 0u && ef_globaltimer() - t0 > 20000000000ull) break;
                __nanosleep(20);
                d[k] = ld_desc(desc + j);
              }
            }
            if (__any_sync(0xFFFFFFFFu, (d[k] >> 62) == 0)) { ok = false; hit = true; }
            const uint32_t inc = __ballot_sync(0xFFFFFFFFu, (d[k] >> 62) == 2);
            const uint32_t first = inc ? (uint32_t)(__ffs((int)inc) - 1) : 31u;
            unsigned long long c = (lane <= first) ? (d[k] & EF_VALMASK) : 0ull;
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o2);
            gex += c;
            if (inc) hit = true;
          }
        }
        if (hit) break;
        idx -= 256;
      }
      if (lane == 0) {
        if (!ok) atomicExch(&ctl->error, 1u);
        st_desc(desc + grp, EF_FLAG_INC | (gex + gsum));
        if (grp == ngroups - 1) ctl->total_out = gex + gsum;
      }
      bases[par * 32u + lane] = out_off + gex + (unsigned long long)(inc_s - tv);
      __threadfence_block();
      asm volatile("bar.arrive %0, %1;" ::"r"(3u + par), "r"(bar_n) : "memory");
    }
    return;
  }

  // ================================================================= workers
  uint32_t prev_total = 0xFFFFFFFFu;                  // tile whose staging window has not left yet
  uint32_t max_recs = 0;                              // most template records a tile of this warp had
  for (uint32_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x, par ^= 1u) {
    const uint32_t tile = grp * nwork + warp;
    const bool active = tile < ntiles;
    const size_t tbase = (size_t)tile * V3_TILE;
    const uint32_t tlen = active ? (uint32_t)((n_eff - tbase < V3_TILE) ? (n_eff - tbase) : V3_TILE) : 0u;
    const bool full = (tlen == V3_TILE);
    const uint32_t lo = lane * 32u;
    const uint32_t cnt_pos = (lo < tlen) ? ((tlen - lo < 32u) ? (tlen - lo) : 32u) : 0u;
    uint32_t w[8];
    if (full) {
      ld_stream32(in + tbase + lo, w);
    } else {
      // 16-byte pieces that start inside the tile (the input buffer is 16-byte granular)
      uint4 v0 = make_uint4(0, 0, 0, 0), v1 = v0;
      if (lo < tlen) v0 = *(const uint4 *)(in + tbase + lo);
      if (lo + 16u < tlen) v1 = *(const uint4 *)(in + tbase + lo + 16u);
      w[0] = v0.x; w[1] = v0.y; w[2] = v0.z; w[3] = v0.w;
      w[4] = v1.x; w[5] = v1.y; w[6] = v1.z; w[7] = v1.w;
    }
    // start states of the two halves
    uint32_t sA = Q, sB = Q;
    if (cnt_pos) {
      const uint32_t cs = chunk_start[tile / V3_TPC];
      const uint32_t smp = *(const uint32_t *)(samples + (size_t)tile * V3_SPT + 2u * lane);
      const uint32_t bp = blockpre[(size_t)tile * (V3_TILE / V3_BLK) + (lane >> 2)];     // 4 lanes per 128-byte block
      const uint32_t sblk = __ldg(F.applyF + (size_t)bp * Q1 + cs);
      sA = __ldg(F.applyF + (size_t)(smp & 0xFFFFu) * Q1 + sblk);
      if (cnt_pos > 16u) sB = __ldg(F.applyF + (size_t)(smp >> 16) * Q1 + sblk);
    }
    const uint32_t lam_tile = (REGS && active) ? lam_end[tile] : 0u;

    // ---- forward walk, live set after each half, count.
    // Exact: the forward walk also multiplies up the backward element of each
    // half; a suffix composition over the warp gives the live set at every half's
    // end.  Speculative (full tiles, programs with registers): the live set at a
    // half's end is guessed from the state there (`guess`, learnt from exactly
    // evaluated tiles), the forward walk skips the backward elements, and the
    // count pass -- which walks the live sets backwards exactly -- must reproduce
    // at every half's start the value assumed for the end of the half before it
    // (by induction from the tile's known end value every guess was then right).
    uint32_t ap[8];
    uint32_t ELA = 0, ELB = 0, accA = 0, accB = 0;
    const uint32_t E0 = (trans_abs + sA * C * STRIDE + slot4) | ((trans_abs + sB * C * STRIDE + slot4) << 16);
    bool spec = false;
    uint32_t gix = 0;                                // guess slots of the two halves' end states
    if (REGS) {
      gix = sB | (__shfl_down_sync(0xFFFFFFFFu, sA, 1) << 16);
      spec = full && (int)spec_ctr >= 0;
      if (spec) {
        const uint32_t gH = guess_ld(guess_abs + (gix & 0xFFFFu));
        const uint32_t gT = (lane == 31u) ? lam_tile : guess_ld(guess_abs + (gix >> 16));
        spec = __all_sync(0xFFFFFFFFu, gH != 0xFFu && gT != 0xFFu);
        ELA = (gH * A) << 24;
        ELB = (gT * A) << 24;
      }
      if (spec) {
        uint32_t EA = E0 << 16, EB = E0 & 0xFFFF0000u, MA = 0, MB = 0;
        v3_forward<LOG, false, true>(cls_abs, w, cnt_pos, EA, EB, MA, MB, ap);
        uint32_t ea = ELA, eb = ELB;
        v3_count<LOG>(pbe, ap, ea, eb, accA, accB);
        // half B starts with what half A assumed at its end; the next lane's half A
        // starts with what this lane's half B assumed at its end
        const uint32_t nextA = __shfl_down_sync(0xFFFFFFFFu, ea, 1);
        const bool good = ((eb ^ ELA) >> 24) == 0u && (lane == 31u || ((nextA ^ ELB) >> 24) == 0u);
        spec = __all_sync(0xFFFFFFFFu, good);
        if (!spec) spec_ctr += 0x10000u;             // a miss
      }
    }
    if (!spec) {
      uint32_t EA = E0 << 16, EB = E0 & 0xFFFF0000u, MA = mulB_abs, MB = mulB_abs;
      if (full) v3_forward<LOG, REGS, true>(cls_abs, w, cnt_pos, EA, EB, MA, MB, ap);
      else v3_forward<LOG, REGS, false>(cls_abs, w, cnt_pos, EA, EB, MA, MB, ap);
      uint32_t lamT = 0, lamH = 0;                   // live set after the thread's last byte / after half A
      if (REGS) {
        const uint32_t mbA = (MA - mulB_abs) >> 8, mbB = (MB - mulB_abs) >> 8;
        uint32_t x = compB[mbA * NB + mbB];          // this thread's element; suffix composition inside the warp
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t y = __shfl_down_sync(0xFFFFFFFFu, x, d);
          if (lane + d < 32u) x = compB[x * NB + y];
        }
        uint32_t ex = __shfl_down_sync(0xFFFFFFFFu, x, 1);
        if (lane == 31u) ex = 0;
        lamT = applyB[ex * NL + lam_tile];
        lamH = applyB[mbB * NL + lamT];
        if (full) {
          // learn from an exactly evaluated tile
          guess_st(guess_abs + (gix & 0xFFFFu), lamH);
          if (lane != 31u) guess_st(guess_abs + (gix >> 16), lamT);
        }
      }
      ELA = (lamH * A) << 24;
      ELB = (lamT * A) << 24;
      accA = 0;
      accB = 0;
      uint32_t ea = ELA, eb = ELB;
      v3_count<LOG>(pbe, ap, ea, eb, accA, accB);
    }
    if (REGS) {
      // tiles in the low half, misses in the high half; guessing stops (sign bit) when a quarter of the tiles miss
      ++spec_ctr;
      if ((spec_ctr & 0xFFFFu) == 64u) spec_ctr = ((spec_ctr >> 16) * 4u > 64u) ? 0x80000000u : 0u;
    }
    const uint32_t cntA = accA & 0x3FFFu, cntB = accB & 0x3FFFu, nrA = accA >> 14, nrB = accB >> 14;
    const uint32_t v = (cntA + cntB) | ((nrA + nrB) << 20);
    uint32_t xs = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) xs = v3_scan_step(xs, d);
    const uint32_t tot = __shfl_sync(0xFFFFFFFFu, xs, 31);
    const uint32_t total = tot & 0xFFFFFu, total_recs = tot >> 20;

    // ---- publish the tile total to the scan warp; prefetch this warp's next tile
    if (lane == 0) slots[par * 32u + warp] = total;
    {
      const size_t nt = (size_t)tile + (size_t)gridDim.x * nwork;
      if (nt < ntiles) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(in + nt * V3_TILE + lane * 32u));
        if (lane < 4u) asm volatile("prefetch.global.L2 [%0];" ::"l"(samples + nt * V3_SPT + lane * 16u));
        if (lane == 4u) asm volatile("prefetch.global.L2 [%0];" ::"l"(blockpre + nt * (V3_TILE / V3_BLK)));
        if (REGS && lane == 5u) asm volatile("prefetch.global.L2 [%0];" ::"l"(lam_end + nt));
      }
    }
    __threadfence_block();
    asm volatile("bar.arrive %0, %1;" ::"r"(1u + par), "r"(bar_n) : "memory");
    // ---- the previous tile leaves its staging window only now: its global offset
    // had this tile's load, forward walk and count pass to arrive
    if (prev_total != 0xFFFFFFFFu) {
      ef_wait_bases(3u + (par ^ 1u), bar_n);
      const unsigned long long gb = bases[(par ^ 1u) * 32u + warp];
      if (gb + prev_total > (unsigned long long)out_cap) {
        if (lane == 0) atomicExch(&ctl->overflow, 1u);
      } else {
        v3_stage_out(stage_abs, prev_total, gb, out, lane);
      }
      __syncwarp();
      prev_total = 0xFFFFFFFFu;
    }
    const uint32_t o_end = xs & 0xFFFFFu;                        // bytes up to and including this thread
    const uint32_t rec_excl = (xs >> 20) - (nrA + nrB);
    max_recs = total_recs > max_recs ? total_recs : max_recs;
    if (total + 16u <= stage_bytes && total_recs <= reccap) {
      // ---- write pass: input bytes into the staging window (tile-relative, so it
      // does not need the global offset), one record per template
      const uint32_t oB = stage_abs + o_end, oA = oB - cntB;
      const uint32_t recpA = recs_abs + 8u * rec_excl, recpB = recpA + 8u * nrA;
      v3_write<LOG, LIT>(pbw, ap, w, ELA, ELB, oA, oB, recpA, recpB);
      __syncwarp();
      // ---- templates: one lane per record
      if (V.rmw_words) {
        // merged edge words: m rounds, lane l takes records l m .. l m + m - 1, one per round, so that
        // neighbours (which may merge the same word) never run in the same round
        const uint32_t m = total_recs > 64u ? (total_recs + 31u) >> 5 : 2u;
        for (uint32_t it = 0; it < m; ++it) {
          const uint32_t r = lane * m + it;
          if (r < total_recs) {
            const uint32_t rc = lds_u32_v(recs_abs + 8u * r), rb = lds_u32_v(recs_abs + 8u * r + 4u);
            const uint32_t ent = (((rc >> 18) << 2) - bew_abs) >> LOG;
            v3_template_rmw(pool_abs, V.pool_stride, rc & 0x3FFFFu, lds_u32(tpl2_abs + 4u * ent), rb, V.rmw_words);
          }
          __syncwarp();
        }
      } else {
        for (uint32_t r = lane; r < total_recs; r += 32u) {
          const uint32_t rc = lds_u32_v(recs_abs + 8u * r), rb = lds_u32_v(recs_abs + 8u * r + 4u);
          const uint32_t ent = (((rc >> 18) << 2) - bew_abs) >> LOG;
          if (F.max_emit <= 3u) v3_copy_short(pool_abs, rc & 0x3FFFFu, lds_u32(tpl2_abs + 4u * ent), rb);
          else v3_copy_template(pool_abs, V.pool_stride, rc & 0x3FFFFu, lds_u32(tpl2_abs + 4u * ent), rb);
        }
        __syncwarp();
      }
      prev_total = total;                                          // staged out after the next tile's count
    } else {
      // ---- the tile's output exceeds the staging window: byte stores to global.
      // Input words and action ids are parked in the (idle) staging window so
      // that the loop can stay rolled.
      ef_wait_bases(3u + par, bar_n);
      const unsigned long long gbase = bases[par * 32u + warp];
      if (gbase + total > (unsigned long long)out_cap) {
        if (lane == 0) atomicExch(&ctl->overflow, 1u);
        continue;
      }
      const uint32_t sp = stage_abs + lane * 64u;
#pragma unroll
      for (int k = 0; k < 8; ++k) { sts_u32(sp + 4u * k, w[k]); sts_u32(sp + 32u + 4u * k, ap[k]); }
      __syncwarp();
      uint8_t *g = out + gbase;
      uint32_t o = o_end, EL = ELB;
      for (int j = 31; j >= 0; --j) {
        const uint32_t a = lds_u8_v(sp + 32u + (uint32_t)j), b = lds_u8_v(sp + (uint32_t)j);
        const uint32_t addr = (pbw + (a << LOG)) + (EL >> (24 - LOG));
        EL = lds_u32(addr);
        o -= EL & 0xFFu;
        if (EL & 0x10000u) g[o] = (uint8_t)b;
        if (LIT && (EL & 0x7F00u)) g[o] = (uint8_t)((EL >> 8) & 0x7Fu);
        if (EL & 0x8000u) {
          const uint32_t t = lds_u32(tpl2_abs + 4u * ((addr - slot4 - bew_abs) >> LOG));
          const uint32_t src = t & 0xFFFFu, tl = (t >> 16) & 0xFFu, hole = t >> 24;
          for (uint32_t k = 0; k < tl; ++k) g[o + k] = (k + 1u == hole) ? (uint8_t)b : (uint8_t)lds_u8(pool_abs + src + k);
        }
      }
      __syncwarp();
    }
  }
  if (lane == 0 && max_recs) atomicMax(&ctl->pad, max_recs);
  if (prev_total != 0xFFFFFFFFu) {
    // the last tile of this warp (par has advanced past its round)
    ef_wait_bases(3u + (par ^ 1u), bar_n);
    const unsigned long long gb = bases[(par ^ 1u) * 32u + warp];
    if (gb + prev_total > (unsigned long long)out_cap) {
      if (lane == 0) atomicExch(&ctl->overflow, 1u);
    } else {
      v3_stage_out(stage_abs, prev_total, gb, out, lane);
    }
  }
}
