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
  unsigned int pad;              // v3: most template records any tile had
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
// byte k of w via one PRMT
__device__ __forceinline__ uint32_t byte_prmt(uint32_t w, int k) { return __byte_perm(w, 0u, 0x4440u | (uint32_t)k); }
// replace byte k of acc by byte 2 of e
__device__ __forceinline__ uint32_t put_byte2(uint32_t acc, uint32_t e, int k) {
  const uint32_t sel = k == 0 ? 0x3216u : k == 1 ? 0x3260u : k == 2 ? 0x3610u : 0x6210u;
  return __byte_perm(acc, e, sel);
}
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

__device__ __forceinline__ unsigned long long ef_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned long long ld_desc(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_desc(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Dynamic shared memory of k_emit_fast.  Everything is addressed as an offset
// from this symbol so that the compiler keeps the accesses in the shared
// address space (LDS/STS with 32-bit addresses) instead of generic loads.
extern __shared__ __align__(128) uint8_t smem_ef[];

struct EmitSmem {                 // byte offsets into smem_ef
  uint32_t cls4;                  // [256]   4 * class
  uint32_t trans2;                // u32 [(Q+1)*C]
  uint32_t BE;                    // u32 [NL*A]
  uint32_t mulB, compB, applyB;   // backward monoid
  uint32_t tplinfo;               // u32 [2*NT]
  uint32_t pool;
  uint32_t in[1];                 // EF_TILE input tile (TMA destination)
  uint32_t stage;                 // staging window
  uint32_t recs;                  // u32 [EF_RECCAP] template records
  uint32_t bar;                   // u64 mbarrier of the input tile
};
#define SM8(off) (smem_ef[(off)])
#define SM16(off) (*(uint16_t *)(smem_ef + (off)))
#define SM32(off) (*(uint32_t *)(smem_ef + (off)))

// forward walk of one thread's 32 bytes: action ids packed 4 per word, backward
// monoid element of the sub-chunk
// In shared memory the low 16 bits of a trans2 entry hold the byte offset of
// the next state's row (so a step is LOP + IADD + LDS), bits 16-23 the action,
// bits 24-31 twice the backward generator; mulB is u16 [NB][NG] holding the
// byte offset of the product's row (so a step is LEA.HI + LDS).
template <bool FULL, bool REGS>
__device__ __forceinline__ void ef_forward(const EmitSmem &S, const uint32_t (&w)[8], uint32_t cnt_pos,
                                           uint32_t &srow, uint32_t (&ap)[8], uint32_t &mbrow) {
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if ((j & 3) == 0) ap[j >> 2] = 0;
    if (FULL || (uint32_t)j < cnt_pos) {
      const uint32_t b = byte_prmt(w[j >> 2], j & 3);
      const uint32_t e = SM32(srow + SM8(b));          // cls4 lives at offset 0
      srow = e & 0xFFFFu;
      ap[j >> 2] = put_byte2(ap[j >> 2], e, j & 3);
      if (REGS) mbrow = SM16(mbrow + (e >> 24));
    }
  }
}

// backward walk: bytes emitted per position.  MODE 0: count (returns bytes in
// bits 16.., 2*records in the low bits).  MODE 1: write the copied input bytes
// to the staging window and push one record per template (branch-free: every
// store is predicated).  MODE 2: write everything to global memory byte by
// byte (tiles whose output exceeds the staging window).
template <bool FULL, int MODE>
__device__ __forceinline__ uint32_t ef_backward(const EmitSmem &S, const uint32_t (&w)[8], const uint32_t (&ap)[8],
                                                uint32_t cnt_pos, uint32_t lamoff, uint32_t o, uint32_t recp,
                                                uint8_t *gout) {
  uint32_t acc = 0;
#pragma unroll
  for (int j = 31; j >= 0; --j) {
    if (FULL || (uint32_t)j < cnt_pos) {
      const uint32_t a = byte_prmt(ap[j >> 2], j & 3);
      const uint32_t e = SM32(S.BE + lamoff + a * 4u);
      lamoff = e & 0xFFFCu;
      if (MODE == 0) {
        acc += e & 0x00FF0002u;
      } else if (MODE == 1) {
        o -= byte_prmt(e, 2);
        // record: staging offset | template id << 16 | input byte << 24
        const uint32_t t = __byte_perm(o, e, 0x3710u);
        const uint32_t sel = (j & 3) == 0 ? 0x4210u : (j & 3) == 1 ? 0x5210u : (j & 3) == 2 ? 0x6210u : 0x7210u;
        const uint32_t rec = __byte_perm(t, w[j >> 2], sel);
        if (e & 2u) { SM32(recp) = rec; recp += 4u; }
        if (e & 1u) SM8(o) = (uint8_t)(rec >> 24);
      } else {
        const uint32_t len = (e >> 16) & 0xFFu;
        o -= len;
        const uint32_t b = byte_prmt(w[j >> 2], j & 3);
        if (e & 2u) {
          const uint32_t i0 = SM32(S.tplinfo + 8u * (e >> 24)), hm = SM32(S.tplinfo + 8u * (e >> 24) + 4u);
          const uint32_t tp = S.pool + (i0 & 0xFFFFu);
          for (uint32_t k = 0; k < len; ++k) gout[o + k] = (k < 32u && ((hm >> k) & 1u)) ? (uint8_t)b : SM8(tp + k);
        } else if (e & 1u) {
          gout[o] = (uint8_t)b;
        }
      }
    }
  }
  return acc;
}

// One record: copy template `id` to staging offset o.  The pool is stored four
// times, copy v shifted left by v bytes, so that an unaligned 32-bit read at
// byte x is the aligned word at (x & ~3) of copy (x & 3): the middle of a
// template moves as whole words, at most 3 + 3 edge bytes move one by one.
__device__ __forceinline__ void ef_copy_template(const EmitSmem &S, uint32_t pool_stride, uint32_t rc) {
  const uint32_t o = rc & 0xFFFFu, id = (rc >> 16) & 0xFFu;
  const uint32_t i0 = SM32(S.tplinfo + 8u * id), hm = SM32(S.tplinfo + 8u * id + 4u);
  const uint32_t src = i0 & 0xFFFFu, len = i0 >> 16;
  uint32_t head = (4u - (o & 3u)) & 3u;
  if (head > len) head = len;
  uint32_t k = 0;
  for (; k < head; ++k) SM8(o + k) = SM8(S.pool + src + k);
  const uint32_t x = src + k;                                  // pool byte offset of the first whole word
  const uint32_t wsrc = S.pool + (x & 3u) * pool_stride + (x & ~3u);
  const uint32_t nw = (len - k) >> 2;
  for (uint32_t i = 0; i < nw; ++i) SM32(o + k + 4u * i) = SM32(wsrc + 4u * i);
  k += 4u * nw;
  for (; k < len; ++k) SM8(o + k) = SM8(S.pool + src + k);
  if (hm) {
    const uint8_t b = (uint8_t)(rc >> 24);
    for (uint32_t h = hm; h; h &= h - 1u) SM8(o + (uint32_t)__ffs((int)h) - 1u) = b;
  }
}

template <bool REGS>
__global__ void __launch_bounds__(EF_NT, 8)
k_emit_fast(PhaseDev P, FastDev F, const uint8_t *__restrict__ in, size_t n_eff, uint32_t ntiles,
            const uint16_t *__restrict__ samples, const uint16_t *__restrict__ chunk_start,
            const uint8_t *__restrict__ lam_end, unsigned long long *__restrict__ desc, FastCtl *__restrict__ ctl,
            uint8_t *__restrict__ out, size_t out_cap, uint32_t stage_bytes) {
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t Q1 = P.Q + 1, C = P.C, A = P.A, NL = F.NL, NB = F.NB, NG = F.NG;
  const uint32_t C4 = 4u * C;
  const uint32_t mb_recip = 65536u / (2u * NG) + 1u;
  __shared__ uint32_t s_tile;
  __shared__ unsigned long long s_warp[EF_NT / 32];
  __shared__ unsigned long long s_total, s_gbase;
  __shared__ uint32_t s_wmb[EF_NT / 32];

  // ---- carve shared memory: tables first (their offsets must fit 16 bits),
  // then the input tile (TMA destination, 128-byte aligned) and the staging window
  EmitSmem S;
  uint32_t sp = 0;
  S.cls4 = sp; sp += 256;             // offset 0: the class lookup needs no address arithmetic
  S.trans2 = sp; sp += Q1 * C * 4u;
  S.BE = sp; sp += NL * A * 4u;
  S.tplinfo = sp; sp += F.NT * 8u;
  S.mulB = sp; sp += NB * NG * 2u;
  S.compB = sp; sp += NB * NB;
  S.applyB = sp; sp += NB * NL;
  sp = (sp + 3u) & ~3u;
  const uint32_t pool_stride = (F.pool_len + 7u) & ~3u;     // four shifted copies, each padded by a word
  S.pool = sp; sp += 4u * pool_stride;
  sp = (sp + 127u) & ~127u;
  S.in[0] = sp; sp += EF_TILE;
  S.stage = sp; sp += stage_bytes + 32u;
  S.recs = sp; sp += EF_RECCAP * 4u;
  S.bar = sp;
  for (uint32_t i = tid; i < Q1 * C; i += EF_NT) {
    const uint32_t e = F.trans2[i];
    SM32(S.trans2 + 4u * i) = (e & 0x00FF0000u) | ((e >> 24) << 25) | (S.trans2 + (e & 0xFFFFu) * C4);
  }
  for (uint32_t i = tid; i < NL * A; i += EF_NT) SM32(S.BE + 4u * i) = F.BE[i];
  for (uint32_t i = tid; i < 2u * F.NT; i += EF_NT) SM32(S.tplinfo + 4u * i) = F.tplinfo[i];
  for (uint32_t i = tid; i < 256; i += EF_NT) SM8(S.cls4 + i) = (uint8_t)(4u * P.cls[i]);
  for (uint32_t i = tid; i < NB * NG; i += EF_NT) SM16(S.mulB + 2u * i) = (uint16_t)(S.mulB + (uint32_t)F.mulB[i] * NG * 2u);
  for (uint32_t i = tid; i < NB * NB; i += EF_NT) SM8(S.compB + i) = F.compB[i];
  for (uint32_t i = tid; i < NB * NL; i += EF_NT) SM8(S.applyB + i) = F.applyB[i];
  for (uint32_t i = tid; i < 4u * pool_stride; i += EF_NT) {
    const uint32_t v = i / pool_stride, k = i - v * pool_stride + v;       // copy v, byte k of the pool
    SM8(S.pool + i) = (k < F.pool_len) ? F.pool[k] : (uint8_t)0;
  }
  uint64_t *bar = (uint64_t *)(smem_ef + S.bar);
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  uint32_t it = 0;
  while (true) {
    // A tile is claimed only when its CTA starts working on it: every tile with
    // a smaller ticket is then finished or in flight, so the chained scan
    // below never waits for work that has not started.
    __syncthreads();
    if (tid == 0) {
      const uint32_t t = atomicAdd(&ctl->ticket, 1u);
      s_tile = t;
      if (t < ntiles) {
        const size_t base = (size_t)t * EF_TILE;
        const uint32_t tl = (uint32_t)((n_eff - base < EF_TILE) ? (n_eff - base) : EF_TILE);
        const uint32_t bytes = (tl + 15u) & ~15u;
        mbar_expect_tx(&bar[0], bytes);
        bulk_g2s(smem_ef + S.in[0], in + base, bytes, &bar[0]);
      }
    }
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tile >= ntiles) break;
    const size_t tbase = (size_t)tile * EF_TILE;
    const uint32_t tlen = (uint32_t)((n_eff - tbase < EF_TILE) ? (n_eff - tbase) : EF_TILE);
    const bool full = (tlen == EF_TILE);
    const uint32_t lo = tid * EF_SUB;
    const uint32_t cnt_pos = (lo < tlen) ? ((tlen - lo < EF_SUB) ? (tlen - lo) : EF_SUB) : 0u;
    // start state of this thread's sub-chunk (independent of the input bytes)
    uint32_t s = P.Q;
    if (cnt_pos) {
      const uint32_t smp = samples[(size_t)tile * EF_NT + tid];
      s = __ldg(F.applyF + (size_t)smp * Q1 + chunk_start[tile]);
    }
    const uint32_t lam_tile = REGS ? lam_end[tile] : 0u;
    mbar_wait(&bar[0], it & 1u);
    uint32_t w[8];
    {
      const uint4 v0 = *(const uint4 *)(smem_ef + S.in[0] + lo), v1 = *(const uint4 *)(smem_ef + S.in[0] + lo + 16);
      w[0] = v0.x; w[1] = v0.y; w[2] = v0.z; w[3] = v0.w;
      w[4] = v1.x; w[5] = v1.y; w[6] = v1.z; w[7] = v1.w;
    }
    // ---- forward walk
    uint32_t ap[8];
    uint32_t srow = S.trans2 + s * C4, mbrow = S.mulB;
    if (full) ef_forward<true, REGS>(S, w, cnt_pos, srow, ap, mbrow);
    else ef_forward<false, REGS>(S, w, cnt_pos, srow, ap, mbrow);
    const uint32_t mb = ((mbrow - S.mulB) * mb_recip) >> 16;      // row offset -> element id (exact, NB <= 255)

    // ---- live set at the end of this thread's sub-chunk
    uint32_t lamoff = 0;
    if (REGS) {
      // inclusive suffix composition inside the warp: x = mb[lane] o mb[lane+1] o ...
      uint32_t x = mb;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_down_sync(0xFFFFFFFFu, x, d);
        if (lane + d < 32u) x = SM8(S.compB + x * NB + y);
      }
      uint32_t ex = __shfl_down_sync(0xFFFFFFFFu, x, 1);     // elements of the later lanes
      if (lane == 31u) ex = 0;
      if (lane == 0) s_wmb[warp] = x;
      __syncthreads();
      for (uint32_t w2 = warp + 1; w2 < EF_NT / 32; ++w2) ex = SM8(S.compB + ex * NB + s_wmb[w2]);
      lamoff = (uint32_t)SM8(S.applyB + ex * NL + lam_tile) * A * 4u;
    }

    // ---- count
    const uint32_t acc = full ? ef_backward<true, 0>(S, w, ap, cnt_pos, lamoff, 0, 0, nullptr)
                              : ef_backward<false, 0>(S, w, ap, cnt_pos, lamoff, 0, 0, nullptr);
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
    if (warp == 0) {
      // ---- tile total, then the chained scan of the tile totals (decoupled look-back)
      unsigned long long t = 0;
      for (uint32_t k = 0; k < EF_NT / 32; ++k) t += s_warp[k];
      const uint32_t total = (uint32_t)t;
      if (lane == 0) st_desc(desc + tile, EF_FLAG_AGG | (unsigned long long)total);
      unsigned long long excl = 0;
      long long idx = (long long)tile - 1;
      bool ok = true;
      while (idx >= 0) {
        const long long j = idx - lane;
        unsigned long long d = EF_FLAG_INC;
        if (j >= 0) {
          const unsigned long long t0 = ef_globaltimer();           // give up only after 20 s of wall time
          uint32_t spins = 0;
          while (((d = ld_desc(desc + j)) >> 62) == 0) {
            if ((++spins & 1023u) == 0u && ef_globaltimer() - t0 > 20000000000ull) break;
            __nanosleep(20);
          }
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
        s_total = t;
        if (tile == ntiles - 1) ctl->total_out = excl + total;
        bulk_wait_read0();              // the previous tile's bulk store has read the staging window
      }
    }
    __syncthreads();
    unsigned long long wbase = 0;
    for (uint32_t k = 0; k < warp; ++k) wbase += s_warp[k];
    const unsigned long long incl = wbase + xs;
    const uint32_t o_end = (uint32_t)incl;                       // bytes up to and including this thread
    const uint32_t rec_base = (uint32_t)((incl - v) >> 32);
    const uint32_t total = (uint32_t)s_total, total_recs = (uint32_t)(s_total >> 32);
    const unsigned long long gbase = s_gbase;
    const uint32_t shift = (uint32_t)(gbase & 15ull);
    const bool fits = gbase + total <= (unsigned long long)out_cap;
    if (!fits) {
      if (tid == 0) atomicExch(&ctl->overflow, 1u);
    } else if (total + shift <= stage_bytes && total_recs <= EF_RECCAP) {
      // ---- write pass into the staging window
      const uint32_t recp = S.recs + 4u * rec_base;
      if (full) ef_backward<true, 1>(S, w, ap, cnt_pos, lamoff, S.stage + shift + o_end, recp, nullptr);
      else ef_backward<false, 1>(S, w, ap, cnt_pos, lamoff, S.stage + shift + o_end, recp, nullptr);
      __syncthreads();
      // ---- templates: one lane per record
      for (uint32_t r = tid; r < total_recs; r += EF_NT) ef_copy_template(S, pool_stride, SM32(S.recs + 4u * r));
      fence_async_smem();
      __syncthreads();
      // ---- staging window -> global: 16-byte words aligned to the destination
      uint8_t *gal = out + (gbase - shift);
      const uint8_t *stg = smem_ef + S.stage;
      const uint32_t end = shift + total;
      const uint32_t w_lo = shift ? 1u : 0u;
      const uint32_t w_hi = end >> 4;
      if (tid == 0 && w_hi > w_lo) {
        bulk_s2g(gal + 16u * w_lo, stg + 16u * w_lo, 16u * (w_hi - w_lo));
        bulk_commit();
      }
      if (shift) {
        const uint32_t he = (end < 16u) ? end : 16u;
        for (uint32_t b2 = shift + tid; b2 < he; b2 += EF_NT) gal[b2] = stg[b2];
      }
      if (w_hi >= w_lo) {
        for (uint32_t b2 = (w_hi << 4) + tid; b2 < end; b2 += EF_NT)
          if (b2 >= shift) gal[b2] = stg[b2];
      }
    } else {
      // ---- output of this tile exceeds the staging window: direct byte stores
      uint8_t *g = out + gbase;
      if (full) ef_backward<true, 2>(S, w, ap, cnt_pos, lamoff, o_end, 0, g);
      else ef_backward<false, 2>(S, w, ap, cnt_pos, lamoff, o_end, 0, g);
    }
    ++it;
  }
  if (tid == 0) bulk_wait_read0();
}
