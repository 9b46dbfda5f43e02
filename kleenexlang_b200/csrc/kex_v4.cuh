// kex_v4.cuh -- G-mode emit kernel ("v4").
//
// Same job as k3_emit (outputconst/outputarray/output and the register buffers
// of crt/crt.c:107-283: the tile that creates a byte writes it), same tile
// geometry and the same chained scan, but with the live-set bookkeeping taken
// out of the per-byte passes:
//
//   For most programs the live set (which registers still reach the output,
//   src/KMC/SymbolicSST.hs:122-136) at a position is a function G of the state
//   there.  G is learnt by the host from the exact live sets at tile ends
//   (k3_seams) and folded into a second transition table (fasttab.py
//   build_gmode / kexcuda.cu learn_gmode): entry (q, c) holds the next row, the
//   number of bytes the transition emits under G and what they are (nothing /
//   the input byte / a template).  A transition that is not G-consistent
//   (BE[G[q']][a].lam_before != G[q]) leads to the FAIL row.  If the exact live
//   set at a tile's end equals G[end state] and no lane reaches the FAIL row,
//   then by induction over the tile's transitions the live set at every
//   position is G[state], exactly.  Such a tile needs
//     forward pass   class lookup + transition lookup per byte (2 LDS), the
//                    entry's emission half is kept (2 bytes per input byte) and
//                    length / template count accumulate on the fma pipe
//     write pass     no lookup: input bytes into the swizzled staging window,
//                    one 8-byte record per template
//     templates      one lane per record, whole words from the byte-shifted
//                    pool copies; the first and last word are merged with what
//                    the window already holds (records are handled in rounds
//                    such that neighbours never merge the same word at once)
//     stage-out      16-byte LDS + 16-byte STG aligned to the destination (the
//                    swizzle permutes 16-byte chunks inside a 128-byte row)
//   Any other tile (the last rows of the input, tiles that would overflow the
//   window) is evaluated exactly by v4_slow_count / v4_slow_write from the
//   tables in global memory -- rare, so slow is fine; KEX_V4_EXACT=1 forces it
//   for every tile (tests).  KEX_V4_KNOCK (bits 2/4/8/16 of force_exact: no write
//   pass / no templates / no stage-out / no look-back) is for timing experiments
//   only: the output is wrong.
#pragma once

#define V4_RECCAP 192u
#define V4_TAIL_TILES 256u       // tail evaluation (kexcuda.cu run_phase): tiles at the end of the input with exact live sets
#define V4_MAX_TPL 64u
#define V4_GB_T 0x80u
#define V4_GB_C 0x40u

struct V4Dev {
  uint32_t ok;
  uint32_t o_trans, o_cls, o_apply, o_tpl, o_pool, o_slots, o_warp;   // byte offsets in dynamic smem
  uint32_t log;             // log2 of the transition table's entry stride (7: per-lane copies, 6: two lanes per copy)
  uint32_t pool_stride;
  uint32_t apply_smem;      // element application table staged in shared memory (u8 [NM][Q+1])
  uint32_t rmw;             // every G-mode template is >= 3 bytes long: edges merged by read-modify-write
  const uint32_t *gtab;     // [(Q+1)*C] next | len << 16 | flags << 24   (device copy, rewritten when G is re-learnt)
  const uint8_t *G;         // [Q+1]     live set the table was built for, 0xFF = none
  const uint32_t *tpl;      // [64]      pool offset | len << 16 | (hole offset + 1) << 24
  const uint8_t *apply8;    // [NM*(Q+1)] applyF as bytes
};

#ifdef KEX_EXP_NOSWZ
__device__ __forceinline__ uint32_t swz4(uint32_t a) { return a; }
#else
__device__ __forceinline__ uint32_t swz4(uint32_t a) { return a ^ ((a >> 3) & 0x70u); }
#endif
__device__ __forceinline__ uint32_t mad_hi_u32(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ uint4 lds_v4_v(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long v4_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// ---- forward step: E = last entry (bits 0-15: absolute shared address of the
// current row in this lane's copy, byte 2: bytes emitted, byte 3: flags)
template <int K, int LOG>
__device__ __forceinline__ void v4_fstep(uint32_t cls_abs, uint32_t &E, uint32_t w, uint32_t &accL, uint32_t &accT) {
  const uint32_t c = lds_u8(__dp4a(w, 1u << (8 * K), cls_abs));
  E = lds_u32((E & 0xFFFFu) + (c << LOG));
  accL = __dp4a(E, 0x00010000u, accL);
  accT = mad_hi_u32(E, 2u, accT);
}

// ---- write pass over one half (16 bytes): a = absolute (unswizzled) window
// address of the next output byte, recp = next record slot
template <int H>
__device__ __forceinline__ void v4_write_half(const uint32_t (&w)[8], const uint32_t (&pr)[16], uint32_t a, uint32_t recp) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const uint32_t p = pr[8 * H + (j >> 1)];
    const uint32_t wreg = w[4 * H + (j >> 2)];
    const int k = j & 3;
    const bool hi = (j & 1) != 0;
#ifdef KEX_EXP_BANKPRIV
    const uint32_t addr = (a & 3u) + ((threadIdx.x & 31u) << 2) + 0x8000u;      // timing only: one bank per lane
#else
    const uint32_t addr = swz4(a);
#endif
    const uint32_t b = k == 0 ? wreg : __umulhi(wreg, 1u << (32 - 8 * k));
    if (p & (hi ? 0x40000000u : 0x4000u)) sts_u8(addr, b);
    if (p & (hi ? 0x80000000u : 0x8000u)) {
      const uint32_t r1 = __byte_perm(p, b, hi ? 0x4432u : 0x4410u);      // len, flags, input byte
      asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(recp), "r"(a), "r"(r1) : "memory");
      recp += 8u;
    }
    a = __dp4a(p, hi ? 0x00010000u : 0x00000001u, a);
  }
}

// ---- one template record (read-modify-write edges): the first and the last word are merged with what the
// window holds (one LOP3 select each), the words in between are plain copies -- no per-word mask logic
__device__ __forceinline__ uint32_t v4_bitsel(uint32_t a, uint32_t b, uint32_t m) {       // (a & m) | (b & ~m)
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(r) : "r"(m), "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ void v4_template_rmw(uint32_t pool_abs, uint32_t pool_stride, uint32_t tpl_abs, uint32_t rc0,
                                                uint32_t rc1) {
  const uint32_t len = rc1 & 0xFFu, id = (rc1 >> 8) & 0x3Fu;
  const uint32_t t = lds_u32(tpl_abs + 4u * id);
  const uint32_t src = t & 0xFFFFu, hole = t >> 24;
  const uint32_t o = rc0, oe = o + len;
  const uint32_t x = src + 4u - (o & 3u), v = x & 3u;
  const uint32_t ps = pool_abs + v * pool_stride + (x - v);
  const uint32_t w0 = o & ~3u, nw = ((oe + 3u) >> 2) - (o >> 2);        // >= 1 (len >= 3)
  const uint32_t mlo = 0xFFFFFFFFu << (8u * (o & 3u)), mhi = 0xFFFFFFFFu >> (8u * ((0u - oe) & 3u));
  {
    const uint32_t m = (nw == 1u) ? (mlo & mhi) : mlo;
    const uint32_t da = swz4(w0);
    sts_u32(da, v4_bitsel(lds_u32(ps), lds_u32_v(da), m));
  }
#pragma unroll
  for (uint32_t i = 1; i < 8; ++i)
    if (i + 1u < nw) sts_u32(swz4(w0 + 4u * i), lds_u32(ps + 4u * i));
  for (uint32_t i = 8; i + 1u < nw; ++i) sts_u32(swz4(w0 + 4u * i), lds_u32(ps + 4u * i));
  if (nw > 1u) {
    const uint32_t da = swz4(w0 + 4u * (nw - 1u));
    sts_u32(da, v4_bitsel(lds_u32(ps + 4u * (nw - 1u)), lds_u32_v(da), mhi));
  }
  if (hole) sts_u8(swz4(o + hole - 1u), (rc1 >> 16) & 0xFFu);
}

// ---- one template record, byte stores at the edges (programs with templates shorter than 3 bytes)
__device__ __forceinline__ void v4_template_bytes(uint32_t pool_abs, uint32_t pool_stride, uint32_t tpl_abs, uint32_t rc0,
                                                  uint32_t rc1) {
  const uint32_t len = rc1 & 0xFFu, id = (rc1 >> 8) & 0x3Fu;
  const uint32_t t = lds_u32(tpl_abs + 4u * id);
  const uint32_t src = t & 0xFFFFu, hole = t >> 24;
  const uint32_t o = rc0;
  const uint32_t pb = pool_abs + 4u + src;                  // copy 0 starts with 4 bytes of padding
  uint32_t head = (0u - o) & 3u;
  if (head > len) head = len;
  const uint32_t nw = (len - head) >> 2, tail = (len - head) & 3u;
#pragma unroll
  for (uint32_t k = 0; k < 3; ++k)
    if (k < head) sts_u8(swz4(o + k), lds_u8(pb + k));
  const uint32_t x = src + 4u + head, v = x & 3u;
  const uint32_t ps = pool_abs + v * pool_stride + (x - v);
  for (uint32_t i = 0; i < nw; ++i) sts_u32(swz4(o + head + 4u * i), lds_u32(ps + 4u * i));
  const uint32_t tb = head + 4u * nw;
#pragma unroll
  for (uint32_t k = 0; k < 3; ++k)
    if (k < tail) sts_u8(swz4(o + tb + k), lds_u8(pb + tb + k));
  if (hole) sts_u8(swz4(o + hole - 1u), (rc1 >> 16) & 0xFFu);
}

// ---- staging window -> global, 16-byte loads and stores aligned to the
// destination: destination chunk c holds output bytes [16c - a, 16c - a + 16)
__device__ __forceinline__ void v4_stage_out(uint32_t stage_abs, uint32_t total, unsigned long long gbase,
                                             uint8_t *__restrict__ out, uint32_t lane) {
  const uint32_t a = (uint32_t)(gbase & 15ull);
  uint8_t *gal = out + (gbase - a);
  const uint32_t end = a + total;
  const uint32_t c_lo = a ? 1u : 0u, c_hi = end >> 4;
  if (a == 0u) {
    for (uint32_t c = lane; c < c_hi; c += 32u) *(uint4 *)(gal + 16u * c) = lds_v4_v(swz4(stage_abs + 16u * c));
  } else {
    const uint32_t S = 16u - a, k = S >> 2, sh = (S & 3u) * 8u;
    for (uint32_t c = c_lo + lane; c < c_hi; c += 32u) {
      const uint4 lo = lds_v4_v(swz4(stage_abs + 16u * (c - 1u))), hi = lds_v4_v(swz4(stage_abs + 16u * c));
      uint4 r;
      if (k == 0u)
        r = make_uint4(__funnelshift_r(lo.x, lo.y, sh), __funnelshift_r(lo.y, lo.z, sh), __funnelshift_r(lo.z, lo.w, sh),
                       __funnelshift_r(lo.w, hi.x, sh));
      else if (k == 1u)
        r = make_uint4(__funnelshift_r(lo.y, lo.z, sh), __funnelshift_r(lo.z, lo.w, sh), __funnelshift_r(lo.w, hi.x, sh),
                       __funnelshift_r(hi.x, hi.y, sh));
      else if (k == 2u)
        r = make_uint4(__funnelshift_r(lo.z, lo.w, sh), __funnelshift_r(lo.w, hi.x, sh), __funnelshift_r(hi.x, hi.y, sh),
                       __funnelshift_r(hi.y, hi.z, sh));
      else
        r = make_uint4(__funnelshift_r(lo.w, hi.x, sh), __funnelshift_r(hi.x, hi.y, sh), __funnelshift_r(hi.y, hi.z, sh),
                       __funnelshift_r(hi.z, hi.w, sh));
      *(uint4 *)(gal + 16u * c) = r;
    }
    // head: destination bytes a..15 of chunk 0 (fewer than 16: one step)
    const uint32_t he = (end < 16u) ? end : 16u;
    if (a + lane < he) gal[a + lane] = (uint8_t)lds_u8_v(swz4(stage_abs + lane));
  }
  // tail: after the last whole chunk (fewer than 16 bytes)
  if (c_hi >= c_lo) {
    const uint32_t b2 = (c_hi << 4) + lane;
    if (b2 < end && b2 >= a) gal[b2] = (uint8_t)lds_u8_v(swz4(stage_abs + b2 - a));
  }
}

// ---- exact evaluation of one lane's bytes from the tables in global memory.
// Returns the lane's backward element; acts[] gets the action ids.
struct V4Slow {
  uint32_t cnt, nrec, lam_end;      // bytes and template records of the lane, live set after its last byte
};

__device__ __noinline__ uint32_t v4_slow_walk(const PhaseDev &P, const FastDev &F, const uint8_t *__restrict__ p, uint32_t n,
                                              uint32_t s, uint8_t *acts) {
  uint32_t mb = 0;
  const uint32_t C = P.C, NG = F.NG;
  for (uint32_t j = 0; j < n; ++j) {
    const uint32_t e = __ldg(F.trans2 + s * C + __ldg(P.cls + p[j]));
    s = e & 0xFFFFu;
    acts[j] = (uint8_t)((e >> 16) & 0xFFu);
    mb = __ldg(F.mulB + mb * NG + (e >> 24));
  }
  return mb;
}

// count pass: the warp's lanes cover consecutive 32-byte pieces of the tile
__device__ __noinline__ V4Slow v4_slow_count(const PhaseDev &P, const FastDev &F, const uint8_t *__restrict__ p, uint32_t n,
                                             uint32_t s, uint32_t lam_tile, uint32_t lane) {
  uint8_t acts[32];
  const uint32_t A = P.A, NL = F.NL, NB = F.NB;
  const uint32_t mb = v4_slow_walk(P, F, p, n, s, acts);
  uint32_t x = mb;                                    // suffix composition inside the warp
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t y = __shfl_down_sync(0xFFFFFFFFu, x, d);
    if (lane + d < 32u) x = __ldg(F.compB + x * NB + y);
  }
  uint32_t ex = __shfl_down_sync(0xFFFFFFFFu, x, 1);
  if (lane == 31u) ex = 0;
  V4Slow r;
  r.lam_end = __ldg(F.applyB + ex * NL + lam_tile);
  r.cnt = 0;
  r.nrec = 0;
  uint32_t L = r.lam_end;
  for (uint32_t j = n; j-- > 0;) {
    const uint32_t e = __ldg(F.BE + L * A + acts[j]);
    r.cnt += (e & 3u) ? ((e >> 16) & 0xFFu) : 0u;
    r.nrec += (e >> 1) & 1u;
    L = ((e & 0xFFFCu) >> 2) / A;
  }
  return r;
}

// write pass: o_end = absolute (unswizzled) window address after the lane's last
// byte, rec_end = address after its last record slot; or, with g != nullptr,
// byte stores to global memory at g[o_end - cnt ...) (tiles that do not fit the window)
__device__ __noinline__ void v4_slow_write(const PhaseDev &P, const FastDev &F, const uint8_t *__restrict__ p, uint32_t n,
                                           uint32_t s, uint32_t lam_end, uint32_t o_end, uint32_t rec_end,
                                           uint8_t *__restrict__ g) {
  uint8_t acts[32];
  const uint32_t A = P.A;
  v4_slow_walk(P, F, p, n, s, acts);
  uint32_t L = lam_end, o = o_end, rp = rec_end;
  for (uint32_t j = n; j-- > 0;) {
    const uint32_t e = __ldg(F.BE + L * A + acts[j]);
    const uint32_t typ = e & 3u, len = (e >> 16) & 0xFFu, b = p[j];
    L = ((e & 0xFFFCu) >> 2) / A;
    if (typ == 0u) continue;
    o -= len;
    if (typ & 1u) {
      if (g) g[o] = (uint8_t)b; else sts_u8(swz4(o), b);
    } else if (g) {
      const uint32_t ti = __ldg(F.tplinfo + 2u * (e >> 24)), hm = __ldg(F.tplinfo + 2u * (e >> 24) + 1u);
      for (uint32_t k = 0; k < len; ++k) g[o + k] = (k < 32u && ((hm >> k) & 1u)) ? (uint8_t)b : __ldg(F.pool + (ti & 0xFFFFu) + k);
    } else {
      rp -= 8u;
      asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(rp), "r"(o), "r"(len | ((V4_GB_T | (e >> 24)) << 8) | (b << 16)) : "memory");
    }
  }
}

// ------------------------------------------------------------------ k4_emit
#ifndef KEX_V4_THREADS
#define KEX_V4_THREADS 1024
#endif
// LOG: log2 of the entry stride of the transition table.  7 = one copy of every entry per lane
// (lane l only touches bank l: conflict-free); 6 = lanes l and l+16 share a copy (tables of up to
// twice the size fit the 16-bit shared addresses, at the price of two-way conflicts).
template <bool REGS, int LOG>
__global__ void __launch_bounds__(KEX_V4_THREADS, 1)
k4_emit(PhaseDev P, FastDev F, V4Dev V, const uint8_t *__restrict__ in, size_t n_eff, uint32_t ntiles,
        const uint16_t *__restrict__ samples, const uint16_t *__restrict__ blockpre,
        const uint16_t *__restrict__ chunk_start, const uint8_t *__restrict__ lam_end, uint32_t tail_first,
        unsigned long long *__restrict__ desc, FastCtl *__restrict__ ctl, uint8_t *__restrict__ out, size_t out_cap,
        unsigned long long out_off, uint32_t stage_bytes, uint32_t warp_bytes, uint32_t reccap, uint32_t force_exact) {
  constexpr uint32_t STRIDE = 1u << LOG;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const uint32_t Q = P.Q, Q1 = Q + 1, C = P.C, C1 = C + 1u;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem_v3);
  const uint32_t slot4 = (lane * 4u) & (STRIDE - 1u);
  const uint32_t trans_abs = base + V.o_trans, cls_abs = base + V.o_cls, tpl_abs = base + V.o_tpl, pool_abs = base + V.o_pool;
  const uint32_t row_bytes = C1 * STRIDE;
  if ((trans_abs & 127u) || trans_abs + Q1 * row_bytes > 65536u || V.log != (uint32_t)LOG) __trap();

  // ---- tables (once per CTA).  Row q has C transition entries and, as entry C, G[q].
  for (uint32_t i = tid; i < Q1 * C1 * (STRIDE / 4u); i += blockDim.x) {
    const uint32_t ent = i >> (LOG - 2), s = i & (STRIDE / 4u - 1u), q = ent / C1, c = ent - q * C1;
    uint32_t v;
    if (c == C) {
      v = V.G[q];
    } else {
      const uint32_t e = V.gtab[q * C + c];
      v = (trans_abs + (e & 0xFFFFu) * row_bytes + s * 4u) | (e & 0xFFFF0000u);
    }
    *(uint32_t *)(smem_v3 + V.o_trans + ent * STRIDE + s * 4u) = v;
  }
  for (uint32_t i = tid; i < 256; i += blockDim.x) smem_v3[V.o_cls + i] = P.cls[i];
  if (V.apply_smem)
    for (uint32_t i = tid; i < F.NM * Q1; i += blockDim.x) smem_v3[V.o_apply + i] = V.apply8[i];
  for (uint32_t i = tid; i < V4_MAX_TPL; i += blockDim.x) *(uint32_t *)(smem_v3 + V.o_tpl + 4u * i) = V.tpl[i];
  for (uint32_t i = tid; i < 4u * V.pool_stride; i += blockDim.x) {
    // copy v, byte i holds padded-pool byte i + v; the padded pool starts with 4 zero bytes
    const uint32_t v = i / V.pool_stride, k = i - v * V.pool_stride + v;
    smem_v3[V.o_pool + i] = (k >= 4u && k - 4u < F.pool_len) ? F.pool[k - 4u] : (uint8_t)0;
  }
  if (tid == 0) *(uint32_t *)(smem_v3 + V.o_slots + 124u) = 0u;          // cta_max_recs (below)
  __syncthreads();

  const uint32_t wreg_off = V.o_warp + warp * warp_bytes;        // this warp's staging window, then its records
  const uint32_t stage_abs = base + wreg_off;
  const uint32_t recs_abs = stage_abs + stage_bytes + 128u;      // window + one row of slack
  const uint32_t nwork = nwarp - 1u;
  volatile uint32_t *slots = (volatile uint32_t *)(smem_v3 + V.o_slots);                             // [2][32] tile totals
  volatile unsigned long long *bases = (volatile unsigned long long *)(smem_v3 + V.o_slots + 256u);  // [2][32] output offsets
  const uint32_t ngroups = (ntiles + nwork - 1u) / nwork;
  const uint32_t nfull = (uint32_t)(n_eff / V3_TILE);           // tiles that lie entirely inside the input
  const uint32_t bar_n = blockDim.x;
  uint32_t par = 0;

  if (warp == nwork) {
    // =============================================================== scan warp
    // (same protocol as k3_emit: one descriptor per group of nwork tiles, decoupled look-back
    // over a window of 256 groups per round trip)
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
      unsigned long long gex = 0;
      long long idx = (long long)grp - 1;
      bool ok = true;
      if (force_exact & 16u) idx = -1;
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
              // a predecessor that has not published yet: poll, giving up only after 20 s of wall time
              const unsigned long long t0 = v4_globaltimer();
              uint32_t spins = 0;
              while ((d[k] >> 62) == 0) {
                if ((++spins & 1023u) == 0u && v4_globaltimer() - t0 > 20000000000ull) break;
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
        if (!ok) atomicCAS(&ctl->error, 0u, 1u);
        st_desc(desc + grp, EF_FLAG_INC | (gex + gsum));
        if (grp == ngroups - 1) ctl->total_out = gex + gsum;
      }
      bases[par * 32u + lane] = out_off + gex + (unsigned long long)(inc_s - tv);
      __threadfence_block();
      asm volatile("bar.arrive %0, %1;" ::"r"(3u + par), "r"(bar_n) : "memory");
    }
    return;
  }

  if ((force_exact & 64u) && blockIdx.x == 0 && tid == 0) atomicExch(&ctl->error, 2u);   // tests: a broken induction
  // ================================================================= workers
  uint32_t prev_total = 0xFFFFFFFFu;                  // tile whose staging window has not left yet
  // most template records any tile of this CTA had (sizes the record slots of the next run): a word of
  // shared memory, not a register per warp -- the loop is short of registers, not of LSU slots
  uint32_t *cta_max_recs = (uint32_t *)(smem_v3 + V.o_slots + 124u);        // slots[0][31]: no worker 31
  const uint32_t fail_row = trans_abs + Q * row_bytes + slot4;
  for (uint32_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x, par ^= 1u) {
    const uint32_t tile = grp * nwork + warp;
    const bool active = tile < ntiles;
    const size_t tbase = (size_t)tile * V3_TILE;
    const bool full = tile < nfull;                     // every tile but (at most) the last one of the input
    const uint32_t lo = lane * 32u;
    uint32_t cnt_pos = 32u;                             // bytes of this lane that lie inside the input
    if (!full) {
      const uint32_t tlen = active ? (uint32_t)((n_eff - tbase < V3_TILE) ? (n_eff - tbase) : V3_TILE) : 0u;
      cnt_pos = (lo < tlen) ? ((tlen - lo < 32u) ? (tlen - lo) : 32u) : 0u;
    }
    uint32_t w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) w[k] = 0;
    if (full) ld_stream32(in + tbase + lo, w);
    // start states of the two halves
    uint32_t sA = Q, sB = Q;
    if (cnt_pos) {
      const uint32_t cs = chunk_start[tile / V3_TPC];
      uint32_t smp = *(const uint32_t *)(samples + (size_t)tile * V3_SPT + 2u * lane);
      if (cnt_pos <= 16u) smp &= 0xFFFFu;               // the second half lies past the input: its sample was never written
      const uint32_t bp = blockpre[(size_t)tile * (V3_TILE / V3_BLK) + (lane >> 2)];     // 4 lanes per 128-byte block
      if (V.apply_smem) {
        const uint32_t ap_abs = base + V.o_apply;
        const uint32_t sblk = lds_u8(ap_abs + bp * Q1 + cs);
        sA = lds_u8(ap_abs + (smp & 0xFFFFu) * Q1 + sblk);
        sB = lds_u8(ap_abs + (smp >> 16) * Q1 + sblk);
      } else {
        const uint32_t sblk = __ldg(F.applyF + (size_t)bp * Q1 + cs);
        sA = __ldg(F.applyF + (size_t)(smp & 0xFFFFu) * Q1 + sblk);
        sB = __ldg(F.applyF + (size_t)(smp >> 16) * Q1 + sblk);
      }
    }
    // exact live set at the tile's end: known for the tiles of the tail only (tail_first = 0: for all)
    const bool has_lam = REGS && active && tile >= tail_first;
    uint32_t lam_tile = has_lam ? lam_end[tile] : 0u;

    // ---- forward walk over the G-mode table: emission halves of the entries, lengths, template counts
    uint32_t pr[16];
    uint32_t cntA = 0, cntB = 0, nrA = 0, nrB = 0;
    bool gmode = full && !(force_exact & 1u);
    uint32_t EB_end = fail_row;
    if (gmode) {
      uint32_t EA = trans_abs + sA * row_bytes + slot4, EB = trans_abs + sB * row_bytes + slot4;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        switch (j & 3) {
          case 0: v4_fstep<0, LOG>(cls_abs, EA, w[j >> 2], cntA, nrA); v4_fstep<0, LOG>(cls_abs, EB, w[4 + (j >> 2)], cntB, nrB); break;
          case 1: v4_fstep<1, LOG>(cls_abs, EA, w[j >> 2], cntA, nrA); v4_fstep<1, LOG>(cls_abs, EB, w[4 + (j >> 2)], cntB, nrB); break;
          case 2: v4_fstep<2, LOG>(cls_abs, EA, w[j >> 2], cntA, nrA); v4_fstep<2, LOG>(cls_abs, EB, w[4 + (j >> 2)], cntB, nrB); break;
          default: v4_fstep<3, LOG>(cls_abs, EA, w[j >> 2], cntA, nrA); v4_fstep<3, LOG>(cls_abs, EB, w[4 + (j >> 2)], cntB, nrB); break;
        }
        if (j & 1) {
          pr[j >> 1] = __byte_perm(pr[j >> 1], EA, 0x7610u);
          pr[8 + (j >> 1)] = __byte_perm(pr[8 + (j >> 1)], EB, 0x7610u);
        } else {
          pr[j >> 1] = EA >> 16;
          pr[8 + (j >> 1)] = EB >> 16;
        }
      }
      // the walk must stay off the FAIL row, half A must end where half B starts, and the tile's
      // exact end live set must be the one the table was built for
      bool bad = ((EA & 0xFFFFu) == fail_row) || ((EB & 0xFFFFu) == fail_row) ||
                 ((EA & 0xFFFFu) != trans_abs + sB * row_bytes + slot4);
      if (has_lam && lane == 31u) bad = bad || (lds_u32((EB & 0xFFFFu) + (C << LOG)) != lam_tile);
      gmode = !__any_sync(0xFFFFFFFFu, bad);
      EB_end = EB;
    }
    V4Slow sl;
    sl.lam_end = 0;
    if (!gmode) {
      // tail evaluation: a tile without exact live sets that is not G-consistent, or the anchor tile
      // (the first one of the tail) itself: the induction has no base -- the host repeats the phase
      // with exact live sets for every tile
      if (REGS && active && tail_first && tile <= tail_first && lane == 0) atomicExch(&ctl->error, 2u);
      // exact evaluation from the tables in global memory (rare)
      sl = v4_slow_count(P, F, in + tbase + lo, cnt_pos, sA, lam_tile, lane);
      cntA = sl.cnt; cntB = 0; nrA = sl.nrec; nrB = 0;
      if (active && lane == 0) atomicAdd(&ctl->ticket, 1u);      // tiles evaluated exactly (host: is G still right?)
    }
#ifdef KEX_EXP_PAD_ALU
    {
      uint32_t z = cntA;
#pragma unroll
      for (int i = 0; i < KEX_EXP_PAD_ALU; ++i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(z) : "r"(cntB), "r"(lane));
      if (z == 0xDEADBEEFu) cntA += 1;
    }
#endif
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
        if (lane < 5u) {      // four sectors of samples, one of block prefixes
          const uint16_t *pf = lane < 4u ? samples + nt * V3_SPT + lane * 16u : blockpre + nt * (V3_TILE / V3_BLK);
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
        }
      }
    }
    __threadfence_block();
    asm volatile("bar.arrive %0, %1;" ::"r"(1u + par), "r"(bar_n) : "memory");
    // ---- the previous tile leaves its staging window only now
    if (prev_total != 0xFFFFFFFFu) {
      ef_wait_bases(3u + (par ^ 1u), bar_n);
      const unsigned long long gb = bases[(par ^ 1u) * 32u + warp];
      if (gb + prev_total > (unsigned long long)out_cap) {
        if (lane == 0) atomicExch(&ctl->overflow, 1u);
      } else if (!(force_exact & 8u)) {
        v4_stage_out(stage_abs, prev_total, gb, out, lane);
      }
      __syncwarp();
      prev_total = 0xFFFFFFFFu;
    }
    const uint32_t o_end = xs & 0xFFFFFu;                        // bytes up to and including this lane
    const uint32_t rec_excl = (xs >> 20) - (nrA + nrB);
    if (lane == 0) atomicMax(cta_max_recs, total_recs);
    if (total + 16u <= stage_bytes && total_recs <= reccap) {
      if (gmode && (force_exact & 2u)) {
      } else if (gmode) {
        const uint32_t aA = stage_abs + o_end - cntA - cntB, aB = aA + cntA;
        const uint32_t recpA = recs_abs + 8u * rec_excl, recpB = recpA + 8u * nrA;
        v4_write_half<0>(w, pr, aA, recpA);
        v4_write_half<1>(w, pr, aB, recpB);
      } else {
        v4_slow_write(P, F, in + tbase + lo, cnt_pos, sA, sl.lam_end, stage_abs + o_end, recs_abs + 8u * (rec_excl + nrA), nullptr);
      }
      __syncwarp();
      // ---- templates: one lane per record.  Merging variant: two neighbours may share a word (records
      // two apart never do when every template has >= 3 bytes), so neighbours go in different rounds
      if (force_exact & 4u) {
      } else if (V.rmw) {
        // m rounds, lane l takes records l m .. l m + m - 1, one per round: neighbours (which may merge the
        // same word) always fall into different rounds, two records of one round are at least m >= 2 apart,
        // and ceil(total / 32) rounds keep nearly every lane busy (90 records: 3 rounds of 30 lanes, where
        // even / odd phases of 32 lanes needed 4)
        const uint32_t m = total_recs > 64u ? (total_recs + 31u) >> 5 : 2u;
        for (uint32_t it = 0; it < m; ++it) {
          const uint32_t r = lane * m + it;
          if (r < total_recs) {
            uint32_t rc0, rc1;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(rc0), "=r"(rc1) : "r"(recs_abs + 8u * r) : "memory");
            v4_template_rmw(pool_abs, V.pool_stride, tpl_abs, rc0, rc1);
          }
          __syncwarp();
        }
      } else {
        for (uint32_t r = lane; r < total_recs; r += 32u)
          v4_template_bytes(pool_abs, V.pool_stride, tpl_abs, lds_u32_v(recs_abs + 8u * r), lds_u32_v(recs_abs + 8u * r + 4u));
        __syncwarp();
      }
      prev_total = total;                                          // staged out after the next tile's forward pass
    } else {
      // ---- the tile's output exceeds the staging window or the record slots: exact path, byte stores to global
      ef_wait_bases(3u + par, bar_n);
      const unsigned long long gbase = bases[par * 32u + warp];
      if (gbase + total > (unsigned long long)out_cap) {
        if (lane == 0) atomicExch(&ctl->overflow, 1u);
        continue;
      }
      if (gmode) {
        // before the tail the live set at the end of a G-consistent tile is G[end state] (by induction
        // from the anchor tile; the host repeats the run exactly if that induction breaks anywhere)
        if (REGS && !has_lam) lam_tile = __shfl_sync(0xFFFFFFFFu, lds_u32((EB_end & 0xFFFFu) + (C << LOG)), 31);
        sl = v4_slow_count(P, F, in + tbase + lo, cnt_pos, sA, lam_tile, lane);
      }
      v4_slow_write(P, F, in + tbase + lo, cnt_pos, sA, sl.lam_end, o_end, 0u, out + gbase);
      __syncwarp();
    }
  }
  if (lane == 0) atomicMax(&ctl->pad, *(volatile uint32_t *)cta_max_recs);   // every warp after its own last update
  if (prev_total != 0xFFFFFFFFu) {
    ef_wait_bases(3u + (par ^ 1u), bar_n);
    const unsigned long long gb = bases[(par ^ 1u) * 32u + warp];
    if (gb + prev_total > (unsigned long long)out_cap) {
      if (lane == 0) atomicExch(&ctl->overflow, 1u);
    } else {
      v4_stage_out(stage_abs, prev_total, gb, out, lane);
    }
  }
}
