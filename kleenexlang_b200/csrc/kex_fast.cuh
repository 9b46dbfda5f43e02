// kex_fast.cuh -- monoid kernels: every pass is one dependent table lookup per
// input byte (tables: kleenexlang_b200/fasttab.py).
//
//   k_fwd_monoid   per 4 KiB chunk (one thread): element of the forward
//                  transition monoid of the chunk, prefix element before every
//                  32-byte sub-chunk.  Replaces the state chain of `matchN`
//                  (src/KMC/Program/Backends/C.hs:72-83) run for all start
//                  states at once.
//   k_seams        per chunk, from its true start state: backward (live-set)
//                  monoid element until it becomes a constant map; exact
//                  position of a failing transition (C.hs:79-81).
//   k_compose_rev / k_push_lam   backward scan: live set at every chunk end.
//   k_emit_fast    per 4 KiB tile (one CTA, 128 threads x 32 B): forward walk
//                  -> action per byte, suffix scan of the threads' backward
//                  elements, backward walk -> bytes emitted per position,
//                  chained scan of the tile totals (decoupled look-back),
//                  backward walk again writing the bytes into a shared-memory
//                  staging window, bulk copy to the output.  Replaces
//                  outputconst/outputarray/output and the register buffers of
//                  crt/crt.c:161-283: the tile that creates a byte writes it.
#pragma once

struct FastDev {
  uint32_t NM, NL, NG, NB, NT, pool_len, max_emit, tbl_bytes;
  const uint16_t *mulF;      // [NM*C]
  const uint16_t *applyF;    // [NM*(Q+1)]
  const uint32_t *trans2;    // [(Q+1)*C]  next | action << 16 | backward generator << 24
  const uint32_t *BE;        // [NL*A]  (lam_before*A*4) | sym | tpl<<1 | len << 16 | x << 24
  const uint8_t *lam_final;  // [Q+1]
  const uint8_t *mulB;       // [NB*NG]
  const uint8_t *compB;      // [NB*NB]
  const uint8_t *applyB;     // [NB*NL]
  const uint8_t *constB;     // [NB]
  const uint32_t *tplinfo;   // [2*NT]  pool offset | len << 16 ; hole mask
  const uint8_t *pool;
};

struct FastCtl {                 // device control block of one emit launch
  unsigned int ticket;
  unsigned int error;            // look-back gave up (should never happen)
  unsigned int overflow;         // output would exceed out_cap
  unsigned int pad;
  unsigned long long total_out;
};

#define EF_NT 128u
#define EF_SUB 32u
#define EF_TILE 4096u
#define EF_RECCAP 1024u
#define EF_FLAG_AGG (1ull << 62)
#define EF_FLAG_INC (2ull << 62)
#define EF_VALMASK ((1ull << 62) - 1)

__device__ __forceinline__ uint32_t byte_at(uint32_t w, int k) { return (w >> (8 * k)) & 0xFFu; }
__device__ __forceinline__ uint32_t word_of(const uint4 &v, int k) {
  return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w;
}

// ------------------------------------------------------------------ K_A
// cls2[b] = 2 * class;  mul[] entries are element ids; address arithmetic is
// in bytes so that the chain is LDS.U16 -> IMAD -> LDS.U16.
#define FM_STEP(word, k)                                                                   \
  m = *(const uint16_t *)((const uint8_t *)mul + m * C2 + cls2[byte_at(word, k)]);
#define FM_STEP16(v)                                                                       \
  FM_STEP(v.x, 0) FM_STEP(v.x, 1) FM_STEP(v.x, 2) FM_STEP(v.x, 3)                          \
  FM_STEP(v.y, 0) FM_STEP(v.y, 1) FM_STEP(v.y, 2) FM_STEP(v.y, 3)                          \
  FM_STEP(v.z, 0) FM_STEP(v.z, 1) FM_STEP(v.z, 2) FM_STEP(v.z, 3)                          \
  FM_STEP(v.w, 0) FM_STEP(v.w, 1) FM_STEP(v.w, 2) FM_STEP(v.w, 3)

__global__ void __launch_bounds__(128, 8)
k_fwd_monoid(PhaseDev P, FastDev F, const uint8_t *__restrict__ in, size_t n, size_t nchunks,
             uint16_t *__restrict__ samples, uint16_t *__restrict__ maps) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint8_t *cls2 = smem;
  uint16_t *mul = (uint16_t *)(smem + 256);
  const uint32_t C = P.C, Q1 = P.Q + 1, C2 = 2u * C;
  for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) cls2[i] = (uint8_t)(2u * P.cls[i]);
  for (uint32_t i = threadIdx.x; i < F.NM * C; i += blockDim.x) mul[i] = F.mulF[i];
  __syncthreads();
  const size_t chunk = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (chunk >= nchunks) return;
  const size_t base = chunk * EF_TILE;
  const uint32_t len = (uint32_t)((n - base < EF_TILE) ? (n - base) : EF_TILE);
  const uint8_t *p = in + base;
  uint16_t *srow = samples + chunk * EF_NT;
  uint32_t m = 0;
  if (len == EF_TILE) {
    for (uint32_t blk = 0; blk < EF_TILE / 256u; ++blk) {
      uint32_t pk[4];
#pragma unroll
      for (int h = 0; h < 4; ++h) {            // 64 bytes: two samples
        const uint8_t *q = p + blk * 256u + h * 64;
        const uint4 v0 = ld_stream16(q), v1 = ld_stream16(q + 16), v2 = ld_stream16(q + 32),
                    v3 = ld_stream16(q + 48);
        pk[h] = m;
        FM_STEP16(v0) FM_STEP16(v1)
        pk[h] |= m << 16;
        FM_STEP16(v2) FM_STEP16(v3)
      }
      *(uint4 *)(srow + blk * 8u) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  } else {
    for (uint32_t j = 0; j < len; ++j) {
      if ((j & (EF_SUB - 1u)) == 0) srow[j / EF_SUB] = (uint16_t)m;
      m = *(const uint16_t *)((const uint8_t *)mul + m * C2 + cls2[p[j]]);
    }
  }
  uint16_t *row = maps + chunk * Q1;
  const uint16_t *ap = F.applyF + (size_t)m * Q1;
  for (uint32_t q = 0; q < Q1; ++q) row[q] = ap[q];
}

// ------------------------------------------------------------------ K_C
__global__ void __launch_bounds__(128)
k_seams(PhaseDev P, FastDev F, const uint8_t *__restrict__ in, size_t n, size_t nchunks,
        const uint16_t *__restrict__ start, const uint16_t *__restrict__ maps,
        uint8_t *__restrict__ bmaps, uint32_t *__restrict__ chunk_fail, RunResult *__restrict__ res) {
  const size_t chunk = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (chunk >= nchunks) return;
  const uint32_t Q = P.Q, C = P.C, NL = F.NL;
  uint32_t s = start[chunk];
  const uint32_t endst = (s == Q) ? Q : maps[chunk * (Q + 1) + s];
  if (chunk == nchunks - 1) res->end_state = endst;
  const bool failing = (s != Q) && (endst == Q);
  uint32_t mb = 0, fail = KEX_NONE32;
  if (s != Q && (failing || NL > 1)) {
    const size_t base = chunk * EF_TILE;
    const uint32_t len = (uint32_t)((n - base < EF_TILE) ? (n - base) : EF_TILE);
    const uint8_t *p = in + base;
    for (uint32_t j = 0; j < len; ++j) {
      const uint32_t e = __ldg(F.trans2 + s * C + __ldg(P.cls + p[j]));
      const uint32_t ns = e & 0xFFFFu;
      if (ns == Q) { fail = j; break; }
      s = ns;
      mb = __ldg(F.mulB + mb * F.NG + (e >> 24));
      if (!failing && __ldg(F.constB + mb)) break;
    }
  }
  chunk_fail[chunk] = fail;
  if (NL > 1)
    for (uint32_t l = 0; l < NL; ++l) bmaps[chunk * NL + l] = __ldg(F.applyB + mb * NL + l);
}

// parent[g] = child[g*G] o child[g*G+1] o ...   (the LAST child is applied first)
__global__ void k_compose_rev(const uint8_t *__restrict__ child, size_t nchild, uint8_t *__restrict__ parent,
                              uint32_t D) {
  const size_t g = blockIdx.x;
  const size_t lo = g * KEX_FANIN;
  const size_t hi = (lo + KEX_FANIN < nchild) ? (lo + KEX_FANIN) : nchild;
  for (uint32_t q = threadIdx.x; q < D; q += blockDim.x) {
    uint32_t s = q;
    for (size_t j = hi; j-- > lo;) s = child[j * D + s];
    parent[g * D + q] = (uint8_t)s;
  }
}

// lam at the END of every child, from lam at the end of its group
__global__ void k_push_lam(const uint8_t *__restrict__ child_maps, size_t nchild,
                           const uint8_t *__restrict__ parent_lam, size_t nparent,
                           uint8_t *__restrict__ child_lam, uint32_t D) {
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nparent) return;
  const size_t lo = g * KEX_FANIN;
  const size_t hi = (lo + KEX_FANIN < nchild) ? (lo + KEX_FANIN) : nchild;
  uint32_t L = parent_lam[g];
  for (size_t j = hi; j-- > lo;) {
    child_lam[j] = (uint8_t)L;
    L = child_maps[j * D + L];
  }
}

__global__ void k_set_u8(uint8_t *p, uint32_t v) { *p = (uint8_t)v; }

// ------------------------------------------------------------------ K_D
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity))
    if (++spins > (1u << 26)) __trap();      // a bulk copy that never lands: fail loudly, do not hang
}
// TMA 1-D bulk copy global -> shared, completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// TMA 1-D bulk copy shared -> global
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ unsigned long long ld_desc(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_desc(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct EmitSmem {                 // pointers into the CTA's dynamic shared memory
  uint8_t *cls4;                  // [256]   4 * class
  uint32_t *trans2;               // [(Q+1)*C]
  uint32_t *BE;                   // [NL*A]
  uint8_t *mulB, *compB, *applyB; // backward monoid
  uint32_t *tplinfo;
  uint8_t *pool;
  uint8_t *in[2];                 // 2 x EF_TILE input ring (TMA destination)
  uint8_t *stage;                 // staging window
  uint32_t *recs;                 // [EF_RECCAP] template records
  uint64_t *bar;                  // [2] mbarriers of the input ring
};

// forward walk of one thread's 32 bytes: action ids packed 4 per word, backward
// monoid element of the sub-chunk
template <bool FULL, bool REGS>
__device__ __forceinline__ void ef_forward(const EmitSmem &S, const uint32_t (&w)[8], uint32_t cnt_pos, uint32_t C4,
                                           uint32_t NG, uint32_t &s, uint32_t (&ap)[8], uint32_t &mb) {
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (FULL || (uint32_t)j < cnt_pos) {
      const uint32_t b = byte_at(w[j >> 2], j & 3);
      const uint32_t e = *(const uint32_t *)((const uint8_t *)S.trans2 + s * C4 + S.cls4[b]);
      s = e & 0xFFFFu;
      if ((j & 3) == 0) ap[j >> 2] = 0;
      ap[j >> 2] |= ((e >> 16) & 0xFFu) << (8 * (j & 3));
      if (REGS) mb = S.mulB[mb * NG + (e >> 24)];
    } else if ((j & 3) == 0) {
      ap[j >> 2] = 0;
    }
  }
}

// backward walk: bytes emitted per position.  MODE 0: count (returns bytes in
// bits 16.., 2*records in the low bits).  MODE 1: write single bytes to the
// staging window and push template records.  MODE 2: write everything to
// global memory byte by byte (tiles whose output exceeds the staging window).
template <bool FULL, int MODE>
__device__ __forceinline__ uint32_t ef_backward(const EmitSmem &S, const FastDev &F, const uint32_t (&w)[8],
                                                const uint32_t (&ap)[8], uint32_t cnt_pos, uint32_t lamoff,
                                                uint32_t o, uint32_t *recp, uint8_t *gout) {
  uint32_t acc = 0;
#pragma unroll
  for (int j = 31; j >= 0; --j) {
    if (FULL || (uint32_t)j < cnt_pos) {
      const uint32_t a = byte_at(ap[j >> 2], j & 3);
      const uint32_t e = *(const uint32_t *)((const uint8_t *)S.BE + lamoff + a * 4u);
      lamoff = e & 0xFFFCu;
      if (MODE == 0) {
        acc += e & 0x00FF0002u;
      } else {
        const uint32_t len = (e >> 16) & 0xFFu;
        o -= len;
        const uint32_t b = byte_at(w[j >> 2], j & 3);
        if (e & 2u) {
          if (MODE == 1) {
            *recp++ = o | ((e >> 24) << 16) | (b << 24);
          } else {
            const uint32_t i0 = S.tplinfo[2 * (e >> 24)], hm = S.tplinfo[2 * (e >> 24) + 1];
            const uint8_t *tp = S.pool + (i0 & 0xFFFFu);
            for (uint32_t k = 0; k < len; ++k) gout[o + k] = (k < 32u && ((hm >> k) & 1u)) ? (uint8_t)b : tp[k];
          }
        } else if (len) {
          const uint8_t v = (uint8_t)((e & 1u) ? b : (e >> 24));
          if (MODE == 1) S.stage[o] = v; else gout[o] = v;
        }
      }
    }
  }
  return acc;
}

template <bool REGS>
__global__ void __launch_bounds__(EF_NT, 6)
k_emit_fast(PhaseDev P, FastDev F, const uint8_t *__restrict__ in, size_t n_eff, uint32_t ntiles,
            const uint16_t *__restrict__ samples, const uint16_t *__restrict__ chunk_start,
            const uint8_t *__restrict__ lam_end, unsigned long long *__restrict__ desc, FastCtl *__restrict__ ctl,
            uint8_t *__restrict__ out, size_t out_cap, uint32_t stage_bytes) {
  extern __shared__ __align__(128) uint8_t smem_ef[];
  uint8_t *smem = smem_ef;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t Q1 = P.Q + 1, C = P.C, A = P.A, NL = F.NL, NB = F.NB, NG = F.NG;
  const uint32_t C4 = 4u * C;
  __shared__ uint32_t s_tile[2];
  __shared__ unsigned long long s_warp[EF_NT / 32];
  __shared__ unsigned long long s_total, s_gbase;
  __shared__ uint32_t s_wmb[EF_NT / 32];

  // ---- carve shared memory: input ring first (TMA wants 16-byte alignment; we give 128)
  EmitSmem S;
  uint8_t *sp = smem;
  S.in[0] = sp; sp += EF_TILE;
  S.in[1] = sp; sp += EF_TILE;
  S.stage = sp; sp += stage_bytes + 32u;
  S.recs = (uint32_t *)sp; sp += EF_RECCAP * 4u;
  S.bar = (uint64_t *)sp; sp += 16;
  S.trans2 = (uint32_t *)sp; sp += Q1 * C * 4u;
  S.BE = (uint32_t *)sp; sp += NL * A * 4u;
  S.tplinfo = (uint32_t *)sp; sp += F.NT * 8u;
  S.cls4 = sp; sp += 256;
  S.mulB = sp; sp += NB * NG;
  S.compB = sp; sp += NB * NB;
  S.applyB = sp; sp += NB * NL;
  S.pool = sp;
  for (uint32_t i = tid; i < Q1 * C; i += EF_NT) S.trans2[i] = F.trans2[i];
  for (uint32_t i = tid; i < NL * A; i += EF_NT) S.BE[i] = F.BE[i];
  for (uint32_t i = tid; i < 2u * F.NT; i += EF_NT) S.tplinfo[i] = F.tplinfo[i];
  for (uint32_t i = tid; i < 256; i += EF_NT) S.cls4[i] = (uint8_t)(4u * P.cls[i]);
  for (uint32_t i = tid; i < NB * NG; i += EF_NT) S.mulB[i] = F.mulB[i];
  for (uint32_t i = tid; i < NB * NB; i += EF_NT) S.compB[i] = F.compB[i];
  for (uint32_t i = tid; i < NB * NL; i += EF_NT) S.applyB[i] = F.applyB[i];
  for (uint32_t i = tid; i < F.pool_len; i += EF_NT) S.pool[i] = F.pool[i];

  auto tile_len = [&](uint32_t t) -> uint32_t {
    const size_t base = (size_t)t * EF_TILE;
    return (uint32_t)((n_eff - base < EF_TILE) ? (n_eff - base) : EF_TILE);
  };
  auto issue_load = [&](uint32_t t, uint32_t slot) {   // one thread
    const uint32_t bytes = (tile_len(t) + 15u) & ~15u;
    mbar_expect_tx(&S.bar[slot], bytes);
    bulk_g2s(S.in[slot], in + (size_t)t * EF_TILE, bytes, &S.bar[slot]);
  };

  if (tid == 0) {
    mbar_init(&S.bar[0], 1);
    mbar_init(&S.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t t0 = atomicAdd(&ctl->ticket, 1u);
    s_tile[0] = t0;
    if (t0 < ntiles) issue_load(t0, 0);
  }
  __syncthreads();

  uint32_t it = 0;
  while (true) {
    const uint32_t slot = it & 1u;
    const uint32_t tile = s_tile[slot];
    if (tile >= ntiles) break;
    // the previous tile's bulk store must have finished reading the staging
    // window before anybody writes to it again (waited for below, before the
    // write pass); take the next ticket and prefetch its input now
    if (tid == 0) {
      const uint32_t tn = atomicAdd(&ctl->ticket, 1u);
      s_tile[slot ^ 1u] = tn;
      if (tn < ntiles) issue_load(tn, slot ^ 1u);
    }
    const uint32_t tlen = tile_len(tile);
    const bool full = (tlen == EF_TILE);
    const uint32_t lo = tid * EF_SUB;
    const uint32_t cnt_pos = (lo < tlen) ? ((tlen - lo < EF_SUB) ? (tlen - lo) : EF_SUB) : 0u;
    // start state of this thread's sub-chunk
    uint32_t s = P.Q;
    if (cnt_pos) {
      const uint32_t smp = samples[(size_t)tile * EF_NT + tid];
      s = __ldg(F.applyF + (size_t)smp * Q1 + chunk_start[tile]);
    }
    const uint32_t lam_tile = REGS ? lam_end[tile] : 0u;
    mbar_wait(&S.bar[slot], (it >> 1) & 1u);
    uint32_t w[8];
    {
      const uint4 v0 = *(const uint4 *)(S.in[slot] + lo), v1 = *(const uint4 *)(S.in[slot] + lo + 16);
      w[0] = v0.x; w[1] = v0.y; w[2] = v0.z; w[3] = v0.w;
      w[4] = v1.x; w[5] = v1.y; w[6] = v1.z; w[7] = v1.w;
    }
    // ---- forward walk
    uint32_t ap[8];
    uint32_t mb = 0;
    if (full) ef_forward<true, REGS>(S, w, cnt_pos, C4, NG, s, ap, mb);
    else ef_forward<false, REGS>(S, w, cnt_pos, C4, NG, s, ap, mb);

    // ---- live set at the end of this thread's sub-chunk
    uint32_t lamoff = 0;
    if (REGS) {
      // inclusive suffix composition inside the warp: x = mb[lane] o mb[lane+1] o ...
      uint32_t x = mb;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_down_sync(0xFFFFFFFFu, x, d);
        if (lane + d < 32u) x = S.compB[x * NB + y];
      }
      uint32_t ex = __shfl_down_sync(0xFFFFFFFFu, x, 1);     // elements of the later lanes
      if (lane == 31u) ex = 0;
      if (lane == 0) s_wmb[warp] = x;
      __syncthreads();
      for (uint32_t w2 = warp + 1; w2 < EF_NT / 32; ++w2) ex = S.compB[ex * NB + s_wmb[w2]];
      lamoff = (uint32_t)S.applyB[ex * NL + lam_tile] * A * 4u;
    }

    // ---- count
    const uint32_t acc = full ? ef_backward<true, 0>(S, F, w, ap, cnt_pos, lamoff, 0, nullptr, nullptr)
                              : ef_backward<false, 0>(S, F, w, ap, cnt_pos, lamoff, 0, nullptr, nullptr);
    const uint32_t cnt = acc >> 16, nrec = (acc & 0xFFFFu) >> 1;
    unsigned long long v = (unsigned long long)cnt | ((unsigned long long)nrec << 32);
    unsigned long long xs = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, xs, d);
      if (lane >= (uint32_t)d) xs += y;
    }
    if (lane == 31u) s_warp[warp] = xs;
    __syncthreads();
    if (tid == 0) {
      unsigned long long t = 0;
      for (uint32_t k = 0; k < EF_NT / 32; ++k) { const unsigned long long q = s_warp[k]; s_warp[k] = t; t += q; }
      s_total = t;
    }
    __syncthreads();
    const unsigned long long incl = s_warp[warp] + xs;
    const uint32_t o_end = (uint32_t)incl;                       // bytes up to and including this thread
    const uint32_t rec_base = (uint32_t)((incl - v) >> 32);
    const uint32_t total = (uint32_t)s_total, total_recs = (uint32_t)(s_total >> 32);

    // ---- chained scan of the tile totals (decoupled look-back), warp 0
    if (warp == 0) {
      if (lane == 0) st_desc(desc + tile, EF_FLAG_AGG | (unsigned long long)total);
      unsigned long long excl = 0;
      long long idx = (long long)tile - 1;
      bool ok = true;
      while (idx >= 0 && ok) {
        const long long j = idx - lane;
        unsigned long long d = EF_FLAG_INC;
        if (j >= 0) {
          uint32_t spins = 0;
          do {
            d = ld_desc(desc + j);
            if ((d >> 62) == 0 && ++spins > (1u << 24)) break;
            if ((d >> 62) == 0) __nanosleep(32);
          } while ((d >> 62) == 0);
        }
        if (__any_sync(0xFFFFFFFFu, (d >> 62) == 0)) { ok = false; break; }
        const uint32_t inc = __ballot_sync(0xFFFFFFFFu, (d >> 62) == 2);
        const uint32_t first = inc ? (uint32_t)(__ffs((int)inc) - 1) : 31u;
        unsigned long long c = (lane <= first) ? (d & EF_VALMASK) : 0ull;
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o2);
        excl += c;
        if (inc) break;
        idx -= 32;
      }
      if (lane == 0) {
        if (!ok) atomicExch(&ctl->error, 1u);
        st_desc(desc + tile, EF_FLAG_INC | (excl + total));
        s_gbase = excl;
        if (tile == ntiles - 1) ctl->total_out = excl + total;
        if (it) bulk_wait_read0();      // staging window free again
      }
    }
    __syncthreads();
    const unsigned long long gbase = s_gbase;
    const uint32_t shift = (uint32_t)(gbase & 15ull);
    const bool fits = gbase + total <= (unsigned long long)out_cap;
    if (!fits) {
      if (tid == 0) atomicExch(&ctl->overflow, 1u);
    } else if (total + shift <= stage_bytes && total_recs <= EF_RECCAP) {
      // ---- write pass into the staging window
      uint32_t *recp = S.recs + rec_base;
      if (full) ef_backward<true, 1>(S, F, w, ap, cnt_pos, lamoff, shift + o_end, recp, nullptr);
      else ef_backward<false, 1>(S, F, w, ap, cnt_pos, lamoff, shift + o_end, recp, nullptr);
      __syncthreads();
      // ---- templates: one lane per record
      for (uint32_t r = tid; r < total_recs; r += EF_NT) {
        const uint32_t rc = S.recs[r];
        const uint32_t o = rc & 0xFFFFu, id = (rc >> 16) & 0xFFu, b = rc >> 24;
        const uint32_t i0 = S.tplinfo[2 * id], hm = S.tplinfo[2 * id + 1];
        const uint8_t *tp = S.pool + (i0 & 0xFFFFu);
        const uint32_t len = i0 >> 16;
        uint8_t *dst = S.stage + o;
        for (uint32_t k = 0; k < len; ++k) dst[k] = (k < 32u && ((hm >> k) & 1u)) ? (uint8_t)b : tp[k];
      }
      fence_async_smem();
      __syncthreads();
      // ---- staging window -> global: 16-byte words aligned to the destination
      uint8_t *gal = out + (gbase - shift);
      const uint32_t end = shift + total;
      const uint32_t w_lo = shift ? 1u : 0u;
      const uint32_t w_hi = end >> 4;
      if (tid == 0 && w_hi > w_lo) {
        bulk_s2g(gal + 16u * w_lo, S.stage + 16u * w_lo, 16u * (w_hi - w_lo));
        bulk_commit();
      }
      if (shift) {
        const uint32_t he = (end < 16u) ? end : 16u;
        for (uint32_t b2 = shift + tid; b2 < he; b2 += EF_NT) gal[b2] = S.stage[b2];
      }
      if (w_hi >= w_lo) {
        for (uint32_t b2 = (w_hi << 4) + tid; b2 < end; b2 += EF_NT)
          if (b2 >= shift) gal[b2] = S.stage[b2];
      }
    } else {
      // ---- output of this tile exceeds the staging window: direct byte stores
      uint8_t *g = out + gbase;
      if (full) ef_backward<true, 2>(S, F, w, ap, cnt_pos, lamoff, o_end, nullptr, g);
      else ef_backward<false, 2>(S, F, w, ap, cnt_pos, lamoff, o_end, nullptr, g);
    }
    ++it;
    __syncthreads();
  }
  if (tid == 0) bulk_wait_read0();
}
