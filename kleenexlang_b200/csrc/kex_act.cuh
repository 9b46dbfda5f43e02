// kex_act.cuh -- action-interpreter phase: register actions with real data
// movement (`reg@t`, `!reg`, `[reg <- ..]`), i.e. the unbounded
// append/concat/reset of crt/crt.c:161-259 when the order of the output is NOT
// the order of creation.
//
// Input is the action stream a transducer phase wrote (encoding:
// kleenexlang_b200/frontend/actions.py -- ESC 0 literal ESC, ESC 1 push,
// ESC 2+2r pop r, ESC 3+2r write r); semantics are those of
// src/KMC/Kleenex/Actions.hs:14-58 as `interp` turns them into register
// updates in src/KMC/SymbolicSST/ActionSST.hs:83-104:
//     byte b   slot[h] ++= b          push     h += 1
//     pop r    R[r] := slot[h]; slot[h] := empty; h -= 1
//     write r  slot[h] ++= R[r]; R[r] := empty
// Every update is copyless, so a byte ends up in at most one place of the final
// output.  No data moves between registers here either: each byte is written
// ONCE, by the tile that reads it, at its final position.  That position comes
// from two prefix computations over slot vectors (32 slots: slot h = builder at
// stack height h, slot 31-r = register r):
//   forward   length of every slot's content (summary per tile: where the old
//             content of a slot goes + bytes added; composed over groups of
//             tiles, scanned, pushed back down)
//   backward  position in the final output where every slot's content ENDS
//             (or NOPOS if it never reaches the output); the summary of a tile
//             is the same fate map + how much the tile appended behind the
//             old content
// A byte appended to slot h is then stored at end[h] - 1 by a backward walk of
// the tile, which undoes the tokens: `write r` needs the length R[r] had, kept
// per token by the exact forward walk (wlen[operand position / 2]).
//
// One thread per tile in every pass (tests/act_model.py is the same algorithm
// in Python, checked against the oracle on CPU).
#pragma once

#define ACT_ESC 0xFFu
#define ACT_NSLOT 32
#define ACT_DEAD 0xFFu
#define ACT_NOPOS 0xFFFFFFFFu
#define ACT_GROUP 256u           // tiles composed per group

struct ActCtl {
  uint32_t err;                  // 1 pop on the bottom builder, 2 stack too deep for the slots, 4 register id out of range
  int32_t hmax;
  int32_t hfinal;
  uint32_t total;                // bytes of the bottom builder at the end of the stream
};

// Forward iteration over the tokens of bytes [lo, hi): BYTE(i, v), PUSH(i), POP(i, r), WRITE(i, r).
#define ACT_FWD_TOKENS(in, lo, hi, BYTE, PUSH, POP, WRITE)                                   \
  {                                                                                          \
    bool pe_ = (lo) > 0 && (in)[(lo) - 1] == ACT_ESC;                                        \
    for (size_t i_ = (lo); i_ < (hi); ++i_) {                                                \
      const uint32_t b_ = (in)[i_];                                                          \
      if (pe_) {                                                                             \
        pe_ = false;                                                                         \
        if (b_ == 0u) { BYTE(i_, ACT_ESC) }                                                  \
        else if (b_ == 1u) { PUSH(i_) }                                                      \
        else if (b_ & 1u) { WRITE(i_, (b_ - 3u) >> 1) }                                      \
        else { POP(i_, (b_ - 2u) >> 1) }                                                     \
      } else if (b_ == ACT_ESC) {                                                            \
        pe_ = true;                                                                          \
      } else { BYTE(i_, b_) }                                                                \
    }                                                                                        \
  }

// ---- P0: stack height
__global__ void __launch_bounds__(128)
ka_heights(const uint8_t *__restrict__ in, size_t n, uint32_t tile, size_t ntiles, uint32_t nregs,
           int32_t *__restrict__ delta, int32_t *__restrict__ mn, int32_t *__restrict__ mx, ActCtl *ctl) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntiles) return;
  const size_t lo = t * tile, hi = (lo + tile < n) ? lo + tile : n;
  int32_t h = 0, l = 0, u = 0;
  bool bad = false;
#define A_BYTE(i, v)
#define A_PUSH(i) { ++h; u = h > u ? h : u; }
#define A_POP(i, r) { --h; l = h < l ? h : l; bad |= (r) >= nregs; }
#define A_WRITE(i, r) { bad |= (r) >= nregs; }
  ACT_FWD_TOKENS(in, lo, hi, A_BYTE, A_PUSH, A_POP, A_WRITE)
#undef A_BYTE
#undef A_PUSH
#undef A_POP
#undef A_WRITE
  delta[t] = h; mn[t] = l; mx[t] = u;
  if (bad) atomicOr(&ctl->err, 4u);
}

// one block: height at the start of every tile, validity
__global__ void __launch_bounds__(1024)
ka_height_scan(const int32_t *__restrict__ delta, const int32_t *__restrict__ mn, const int32_t *__restrict__ mx,
               size_t ntiles, uint32_t nregs, int32_t *__restrict__ h0, ActCtl *ctl) {
  __shared__ long long part[1024];
  const size_t seg = (ntiles + blockDim.x - 1) / blockDim.x;
  const size_t lo = (size_t)threadIdx.x * seg < ntiles ? (size_t)threadIdx.x * seg : ntiles;
  const size_t hi = lo + seg < ntiles ? lo + seg : ntiles;
  long long s = 0;
  for (size_t t = lo; t < hi; ++t) s += delta[t];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long acc = 0;
    for (uint32_t k = 0; k < blockDim.x; ++k) { const long long v = part[k]; part[k] = acc; acc += v; }
  }
  __syncthreads();
  long long h = part[threadIdx.x];
  long long top = 0;
  uint32_t err = 0;
  for (size_t t = lo; t < hi; ++t) {
    h0[t] = (int32_t)h;
    if (h + mn[t] < 0) err |= 1u;
    if (h + mx[t] > top) top = h + mx[t];
    h += delta[t];
  }
  if (top + 1 + (long long)nregs > ACT_NSLOT) err |= 2u;
  if (lo < hi && hi == ntiles) { h0[ntiles] = (int32_t)h; ctl->hfinal = (int32_t)h; }
  if (err) atomicOr(&ctl->err, err);
  atomicMax(&ctl->hmax, (int32_t)top);
}

// ---- slot-map primitives (arrays are thread-local)
__device__ __forceinline__ void act_pop(uint8_t *fate, uint32_t h, uint32_t N) {
#pragma unroll 4
  for (int x = 0; x < ACT_NSLOT; ++x) {
    const uint32_t f = fate[x];
    fate[x] = (uint8_t)(f == N ? ACT_DEAD : (f == h ? N : f));
  }
}

// ---- P1: forward summary of a tile (fate of every slot's old content, bytes added per slot)
__global__ void __launch_bounds__(128)
ka_fwd_summary(const uint8_t *__restrict__ in, size_t n, uint32_t tile, size_t ntiles,
               const int32_t *__restrict__ h0, uint8_t *__restrict__ fate_out, uint32_t *__restrict__ add_out) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntiles) return;
  const size_t lo = t * tile, hi = (lo + tile < n) ? lo + tile : n;
  uint8_t fate[ACT_NSLOT];
  uint32_t add[ACT_NSLOT];
  for (int x = 0; x < ACT_NSLOT; ++x) { fate[x] = (uint8_t)x; add[x] = 0; }
  uint32_t h = (uint32_t)h0[t];
#define A_BYTE(i, v) { add[h] += 1u; }
#define A_PUSH(i) { ++h; }
#define A_POP(i, r) { const uint32_t N = ACT_NSLOT - 1u - (r); act_pop(fate, h, N); add[N] = add[h]; add[h] = 0; --h; }
#define A_WRITE(i, r) { const uint32_t N = ACT_NSLOT - 1u - (r);                              \
    for (int x = 0; x < ACT_NSLOT; ++x) if (fate[x] == N) fate[x] = (uint8_t)h;               \
    add[h] += add[N]; add[N] = 0; }
  ACT_FWD_TOKENS(in, lo, hi, A_BYTE, A_PUSH, A_POP, A_WRITE)
#undef A_BYTE
#undef A_PUSH
#undef A_POP
#undef A_WRITE
  for (int x = 0; x < ACT_NSLOT; ++x) { fate_out[t * ACT_NSLOT + x] = fate[x]; add_out[t * ACT_NSLOT + x] = add[x]; }
}

// forward composition of the summaries of one group of tiles (earlier first)
__global__ void __launch_bounds__(128)
ka_group_compose(const uint8_t *__restrict__ fate, const uint32_t *__restrict__ add, size_t ntiles, size_t ngroups,
                 uint8_t *__restrict__ gfate, uint32_t *__restrict__ gadd) {
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ngroups) return;
  const size_t lo = g * ACT_GROUP, hi = (lo + ACT_GROUP < ntiles) ? lo + ACT_GROUP : ntiles;
  uint8_t f[ACT_NSLOT];
  uint32_t a[ACT_NSLOT], b[ACT_NSLOT];
  for (int x = 0; x < ACT_NSLOT; ++x) { f[x] = fate[lo * ACT_NSLOT + x]; a[x] = add[lo * ACT_NSLOT + x]; }
  for (size_t t = lo + 1; t < hi; ++t) {
    const uint8_t *f2 = fate + t * ACT_NSLOT;
    const uint32_t *a2 = add + t * ACT_NSLOT;
    for (int x = 0; x < ACT_NSLOT; ++x) b[x] = a2[x];
    for (int m = 0; m < ACT_NSLOT; ++m) { const uint32_t d = f2[m]; if (d != ACT_DEAD) b[d] += a[m]; }
    for (int x = 0; x < ACT_NSLOT; ++x) { a[x] = b[x]; const uint32_t d = f[x]; f[x] = (uint8_t)(d == ACT_DEAD ? ACT_DEAD : f2[d]); }
  }
  for (int x = 0; x < ACT_NSLOT; ++x) { gfate[g * ACT_NSLOT + x] = f[x]; gadd[g * ACT_NSLOT + x] = a[x]; }
}

// one warp (lane = slot): slot lengths at the start of every group; gvec[ngroups] = at the end of the stream
__global__ void __launch_bounds__(32)
ka_group_scan(const uint8_t *__restrict__ gfate, const uint32_t *__restrict__ gadd, size_t ngroups,
              uint32_t *__restrict__ gvec) {
  const uint32_t lane = threadIdx.x;
  uint32_t v = 0;
  for (size_t g = 0; g < ngroups; ++g) {
    gvec[g * ACT_NSLOT + lane] = v;
    const uint32_t f = gfate[g * ACT_NSLOT + lane];
    uint32_t acc = gadd[g * ACT_NSLOT + lane];
    for (int m = 0; m < ACT_NSLOT; ++m) {
      const uint32_t fm = __shfl_sync(0xFFFFFFFFu, f, m), vm = __shfl_sync(0xFFFFFFFFu, v, m);
      if (fm == lane) acc += vm;
    }
    v = acc;
  }
  gvec[ngroups * ACT_NSLOT + lane] = v;
}

// slot lengths at the start of every tile
__global__ void __launch_bounds__(128)
ka_tile_vectors(const uint8_t *__restrict__ fate, const uint32_t *__restrict__ add, size_t ntiles, size_t ngroups,
                const uint32_t *__restrict__ gvec, uint32_t *__restrict__ vec) {
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ngroups) return;
  const size_t lo = g * ACT_GROUP, hi = (lo + ACT_GROUP < ntiles) ? lo + ACT_GROUP : ntiles;
  uint32_t v[ACT_NSLOT], b[ACT_NSLOT];
  for (int x = 0; x < ACT_NSLOT; ++x) v[x] = gvec[g * ACT_NSLOT + x];
  for (size_t t = lo; t < hi; ++t) {
    for (int x = 0; x < ACT_NSLOT; ++x) { vec[t * ACT_NSLOT + x] = v[x]; b[x] = add[t * ACT_NSLOT + x]; }
    for (int m = 0; m < ACT_NSLOT; ++m) { const uint32_t d = fate[t * ACT_NSLOT + m]; if (d != ACT_DEAD) b[d] += v[m]; }
    for (int x = 0; x < ACT_NSLOT; ++x) v[x] = b[x];
  }
}

// ---- P2: exact lengths; length of R[r] at every `write r`; backward summary of the tile
// (fate as in P1; tail = bytes behind a slot's old content inside the slot it ended up in)
__global__ void __launch_bounds__(128)
ka_fwd_exact(const uint8_t *__restrict__ in, size_t n, uint32_t tile, size_t ntiles, const int32_t *__restrict__ h0,
             const uint32_t *__restrict__ vec, uint32_t *__restrict__ wlen, uint8_t *__restrict__ fate_out,
             uint32_t *__restrict__ tail_out, ActCtl *ctl) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntiles) return;
  const size_t lo = t * tile, hi = (lo + tile < n) ? lo + tile : n;
  uint8_t fate[ACT_NSLOT];
  uint32_t len[ACT_NSLOT], endoff[ACT_NSLOT];
  for (int x = 0; x < ACT_NSLOT; ++x) { fate[x] = (uint8_t)x; len[x] = endoff[x] = vec[t * ACT_NSLOT + x]; }
  uint32_t h = (uint32_t)h0[t];
#define A_BYTE(i, v) { len[h] += 1u; }
#define A_PUSH(i) { ++h; }
#define A_POP(i, r) { const uint32_t N = ACT_NSLOT - 1u - (r); act_pop(fate, h, N); len[N] = len[h]; len[h] = 0; --h; }
#define A_WRITE(i, r) { const uint32_t N = ACT_NSLOT - 1u - (r);                              \
    wlen[(i) >> 1] = len[N];                                                                  \
    for (int x = 0; x < ACT_NSLOT; ++x) if (fate[x] == N) { fate[x] = (uint8_t)h; endoff[x] += len[h]; } \
    len[h] += len[N]; len[N] = 0; }
  ACT_FWD_TOKENS(in, lo, hi, A_BYTE, A_PUSH, A_POP, A_WRITE)
#undef A_BYTE
#undef A_PUSH
#undef A_POP
#undef A_WRITE
  for (int x = 0; x < ACT_NSLOT; ++x) {
    const uint32_t d = fate[x];
    fate_out[t * ACT_NSLOT + x] = (uint8_t)d;
    tail_out[t * ACT_NSLOT + x] = (d == ACT_DEAD) ? 0u : len[d] - endoff[x];
  }
  if (t == ntiles - 1) ctl->total = len[0];
}

// ---- P3: backward composition (earlier tile first, then the later one)
__global__ void __launch_bounds__(128)
ka_group_bcompose(const uint8_t *__restrict__ fate, const uint32_t *__restrict__ tail, size_t ntiles, size_t ngroups,
                  uint8_t *__restrict__ gfate, uint32_t *__restrict__ gtail) {
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ngroups) return;
  const size_t lo = g * ACT_GROUP, hi = (lo + ACT_GROUP < ntiles) ? lo + ACT_GROUP : ntiles;
  uint8_t f[ACT_NSLOT];
  uint32_t o[ACT_NSLOT];
  for (int x = 0; x < ACT_NSLOT; ++x) { f[x] = fate[lo * ACT_NSLOT + x]; o[x] = tail[lo * ACT_NSLOT + x]; }
  for (size_t t = lo + 1; t < hi; ++t) {
    const uint8_t *f2 = fate + t * ACT_NSLOT;
    const uint32_t *o2 = tail + t * ACT_NSLOT;
    for (int x = 0; x < ACT_NSLOT; ++x) {
      const uint32_t d = f[x];
      if (d == ACT_DEAD) continue;
      const uint32_t d2 = f2[d];
      if (d2 == ACT_DEAD) { f[x] = ACT_DEAD; o[x] = 0; }
      else { f[x] = (uint8_t)d2; o[x] += o2[d]; }
    }
  }
  for (int x = 0; x < ACT_NSLOT; ++x) { gfate[g * ACT_NSLOT + x] = f[x]; gtail[g * ACT_NSLOT + x] = o[x]; }
}

// one warp: end position of every slot at the END of every group
__global__ void __launch_bounds__(32)
ka_group_bscan(const uint8_t *__restrict__ gfate, const uint32_t *__restrict__ gtail, size_t ngroups,
               const ActCtl *__restrict__ ctl, uint32_t *__restrict__ gvec) {
  const uint32_t lane = threadIdx.x;
  uint32_t e = (lane == 0) ? ctl->total : ACT_NOPOS;
  for (size_t g = ngroups; g-- > 0;) {
    gvec[g * ACT_NSLOT + lane] = e;
    const uint32_t f = gfate[g * ACT_NSLOT + lane], o = gtail[g * ACT_NSLOT + lane];
    const uint32_t ef = __shfl_sync(0xFFFFFFFFu, e, f & 31u);
    e = (f == ACT_DEAD || ef == ACT_NOPOS) ? ACT_NOPOS : ef - o;
  }
}

// end positions at the end of every tile
__global__ void __launch_bounds__(128)
ka_tile_bvectors(const uint8_t *__restrict__ fate, const uint32_t *__restrict__ tail, size_t ntiles, size_t ngroups,
                 const uint32_t *__restrict__ gvec, uint32_t *__restrict__ vec) {
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ngroups) return;
  const size_t lo = g * ACT_GROUP, hi = (lo + ACT_GROUP < ntiles) ? lo + ACT_GROUP : ntiles;
  uint32_t e[ACT_NSLOT], b[ACT_NSLOT];
  for (int x = 0; x < ACT_NSLOT; ++x) e[x] = gvec[g * ACT_NSLOT + x];
  for (size_t t = hi; t-- > lo;) {
    for (int x = 0; x < ACT_NSLOT; ++x) vec[t * ACT_NSLOT + x] = e[x];
    for (int x = 0; x < ACT_NSLOT; ++x) {
      const uint32_t d = fate[t * ACT_NSLOT + x];
      const uint32_t ed = (d == ACT_DEAD) ? ACT_NOPOS : e[d];
      b[x] = (ed == ACT_NOPOS) ? ACT_NOPOS : ed - tail[t * ACT_NSLOT + x];
    }
    for (int x = 0; x < ACT_NSLOT; ++x) e[x] = b[x];
  }
}

// ---- P4: backward walk of the tile; every surviving byte is stored at its final position
__global__ void __launch_bounds__(128)
ka_write(const uint8_t *__restrict__ in, size_t n, uint32_t tile, size_t ntiles, const int32_t *__restrict__ h0,
         const uint32_t *__restrict__ vec, const uint32_t *__restrict__ wlen, uint8_t *__restrict__ out) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntiles) return;
  const size_t lo = t * tile, hi = (lo + tile < n) ? lo + tile : n;
  uint32_t e[ACT_NSLOT];
  for (int x = 0; x < ACT_NSLOT; ++x) e[x] = vec[t * ACT_NSLOT + x];
  uint32_t h = (uint32_t)h0[t + 1];
  size_t i = hi;
  while (i > lo) {
    --i;
    const uint32_t b = in[i];
    const bool operand = i > 0 && in[i - 1] == ACT_ESC;
    if (operand) {
      if (b == 0u) {
        if (e[h] != ACT_NOPOS) out[--e[h]] = (uint8_t)ACT_ESC;
      } else if (b == 1u) {
        --h;                                          // undo push
      } else if (b & 1u) {                            // undo write r
        const uint32_t N = ACT_NSLOT - 1u - ((b - 3u) >> 1);
        const uint32_t eh = e[h];
        e[N] = eh;
        if (eh != ACT_NOPOS) e[h] = eh - wlen[i >> 1];
      } else {                                        // undo pop r
        const uint32_t N = ACT_NSLOT - 1u - ((b - 2u) >> 1);
        ++h;
        e[h] = e[N];
        e[N] = ACT_NOPOS;
      }
      if (i > lo) --i;                                // the ESC lead
      else break;
    } else if (b != ACT_ESC) {
      if (e[h] != ACT_NOPOS) out[--e[h]] = (uint8_t)b;
    }
  }
}
