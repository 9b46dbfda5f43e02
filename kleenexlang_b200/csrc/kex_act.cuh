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
// One thread per tile in the passes over the stream, one warp (lane = slot) per
// group of tiles in the scans (tests/act_model.py is the same algorithm in
// Python, checked against the oracle on CPU).
#pragma once

#define ACT_ESC 0xFFu
#define ACT_NSLOT 32
#define ACT_DEAD 0xFFu
#define ACT_NOPOS 0xFFFFFFFFu
#define ACT_GROUP 256u           // tiles composed per group (one warp); the scan over groups is sequential

struct ActCtl {
  uint32_t err;                  // 1 pop on the bottom builder, 2 stack too deep for the slots, 4 register id out of range
  int32_t hmax;
  int32_t hfinal;
  uint32_t total;                // bytes of the bottom builder at the end of the stream
};

#define ACT_NT 128u              // threads per block of the per-tile kernels

// Thread-private slot arrays live in shared memory, [slot][thread]: a thread only
// ever touches its own column (bank = thread), and unlike local memory the
// footprint cannot fall out of L1 when many blocks are resident.
#define ACT_S(arr, slot_) arr[(slot_) * ACT_NT + threadIdx.x]

// Four input bytes; the ragged last word of the stream is read byte by byte so
// that nothing behind it is touched (tiles start on 16-byte boundaries of a
// 16-byte aligned stream).
__device__ __forceinline__ uint32_t act_ld4(const uint8_t *__restrict__ in, size_t c, uint32_t cnt) {
  if (cnt == 4u) return __ldg((const uint32_t *)(in + c));
  uint32_t w = 0;
  for (uint32_t k = 0; k < cnt; ++k) w |= (uint32_t)in[c + k] << (8u * k);
  return w;
}

// Forward iteration over the tokens of bytes [lo, hi): BYTE(i, v), PUSH(i), POP(i, r), WRITE(i, r).
// (word loop rolled, four copies of the body: the kernels are bound by instruction fetch otherwise)
#define ACT_FWD_TOKENS(in, lo, hi, BYTE, PUSH, POP, WRITE)                                   \
  {                                                                                          \
    bool pe_ = (lo) > 0 && (in)[(lo) - 1] == ACT_ESC;                                        \
    _Pragma("unroll 1")                                                                      \
    for (size_t c_ = (lo); c_ < (hi); c_ += 4) {                                             \
      const uint32_t cnt_ = ((hi) - c_ < 4) ? (uint32_t)((hi) - c_) : 4u;                    \
      const uint32_t w_ = act_ld4(in, c_, cnt_);                                             \
      _Pragma("unroll")                                                                      \
      for (uint32_t k_ = 0; k_ < 4u; ++k_) {                                                 \
        if (k_ < cnt_) {                                                                     \
          const uint32_t b_ = (w_ >> (8u * k_)) & 0xFFu;                                     \
          const size_t i_ = c_ + k_;                                                         \
          (void)i_;                                                                          \
          if (pe_) {                                                                         \
            pe_ = false;                                                                     \
            if (b_ == 0u) { BYTE(i_, ACT_ESC) }                                              \
            else if (b_ == 1u) { PUSH(i_) }                                                  \
            else if (b_ & 1u) { WRITE(i_, (b_ - 3u) >> 1) }                                  \
            else { POP(i_, (b_ - 2u) >> 1) }                                                 \
          } else if (b_ == ACT_ESC) {                                                        \
            pe_ = true;                                                                      \
          } else { BYTE(i_, b_) }                                                            \
        }                                                                                    \
      }                                                                                      \
    }                                                                                        \
  }

// ---- P0: stack height
// inclusive scan of one value per thread over a block of ACT_NT threads
__device__ __forceinline__ long long act_block_scan(long long v, long long *warp_tot, long long &block_total) {
  const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const long long y = __shfl_up_sync(0xFFFFFFFFu, v, d);
    if (lane >= (uint32_t)d) v += y;
  }
  if (lane == 31u) warp_tot[w] = v;
  __syncthreads();
  long long base = 0, tot = 0;
  for (uint32_t k = 0; k < ACT_NT / 32u; ++k) { if (k < w) base += warp_tot[k]; tot += warp_tot[k]; }
  block_total = tot;
  return v + base;
}

// per tile: net height change, lowest and highest height relative to the tile start; per block: sum of the changes
__global__ void __launch_bounds__(ACT_NT)
ka_heights(const uint8_t *__restrict__ in, size_t n, uint32_t tile, size_t ntiles, uint32_t nregs,
           int32_t *__restrict__ delta, int32_t *__restrict__ mn, int32_t *__restrict__ mx,
           long long *__restrict__ bsum, ActCtl *ctl) {
  __shared__ long long warp_tot[ACT_NT / 32u];
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  int32_t h = 0, l = 0, u = 0;
  bool bad = false;
  if (t < ntiles) {
    const size_t lo = t * tile, hi = (lo + tile < n) ? lo + tile : n;
    // Branch-free over the four bytes of a word (this pass only counts, so the lanes of a warp need not
    // diverge on the token kinds): E = bytes equal to ESC, O = bytes that follow an ESC = operands.
    uint32_t pe = (lo > 0 && in[lo - 1] == ACT_ESC) ? 0x80u : 0u;
#pragma unroll 1
    for (size_t c = lo; c < hi; c += 4) {
      const uint32_t cnt = (hi - c < 4) ? (uint32_t)(hi - c) : 4u;
      const uint32_t w = act_ld4(in, c, cnt);
      const uint32_t x = ~w;                                     // a byte of x is zero iff the byte of w is 0xFF
      const uint32_t E = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);
      const uint32_t O = (E << 8) | pe;
      bad |= (E & O) != 0u;                                      // an operand is never ESC in an action stream
      pe = E >> 24;                                              // 0x80 iff the word's last byte is an ESC
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t b = (w >> (8 * k)) & 0xFFu;
        const int32_t o = (int32_t)((O >> (8 * k + 7)) & 1u);
        const int32_t push = o & (int32_t)(b == 1u);
        const int32_t reg_op = o & (int32_t)(b >= 2u);           // pop (even) or write (odd)
        const int32_t pop = reg_op & (int32_t)((b & 1u) == 0u);
        h += push;
        u = h > u ? h : u;
        h -= pop;
        l = h < l ? h : l;
        bad |= (reg_op != 0) && (((b - 2u) >> 1) >= nregs);
      }
    }
    delta[t] = h; mn[t] = l; mx[t] = u;
    if (bad) atomicOr(&ctl->err, 4u);
  }
  long long tot;
  act_block_scan((long long)h, warp_tot, tot);
  if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

// one block: exclusive scan of the per-block sums (in place); the height at the end of the stream
__global__ void __launch_bounds__(1024)
ka_height_scan(long long *__restrict__ bsum, size_t nblocks, ActCtl *ctl) {
  __shared__ long long part[1024];
  const size_t seg = (nblocks + blockDim.x - 1) / blockDim.x;
  const size_t lo = (size_t)threadIdx.x * seg < nblocks ? (size_t)threadIdx.x * seg : nblocks;
  const size_t hi = lo + seg < nblocks ? lo + seg : nblocks;
  long long s = 0;
  for (size_t b = lo; b < hi; ++b) s += bsum[b];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long acc = 0;
    for (uint32_t k = 0; k < blockDim.x; ++k) { const long long v = part[k]; part[k] = acc; acc += v; }
    ctl->hfinal = (int32_t)acc;
  }
  __syncthreads();
  long long h = part[threadIdx.x];
  for (size_t b = lo; b < hi; ++b) { const long long v = bsum[b]; bsum[b] = h; h += v; }
}

// height at the start of every tile (h0[ntiles] = at the end of the stream); validity
__global__ void __launch_bounds__(ACT_NT)
ka_tile_heights(const int32_t *__restrict__ delta, const int32_t *__restrict__ mn, const int32_t *__restrict__ mx,
                const long long *__restrict__ bsum, size_t ntiles, uint32_t nregs, int32_t *__restrict__ h0, ActCtl *ctl) {
  __shared__ long long warp_tot[ACT_NT / 32u];
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const long long d = (t < ntiles) ? (long long)delta[t] : 0ll;
  long long tot;
  const long long incl = act_block_scan(d, warp_tot, tot);
  if (t >= ntiles) return;
  const long long h = bsum[blockIdx.x] + incl - d;
  h0[t] = (int32_t)h;
  if (t == ntiles - 1) h0[ntiles] = (int32_t)(h + d);
  uint32_t err = 0;
  if (h + mn[t] < 0) err |= 1u;
  if (h + mx[t] + 1 + (long long)nregs > ACT_NSLOT) err |= 2u;
  if (err) atomicOr(&ctl->err, err);
  atomicMax(&ctl->hmax, (int32_t)(h + mx[t]));
}

// ---- slot-content bookkeeping of a tile
// Which slot the OLD content of every slot (its content at the tile start) is in,
// and where it ends inside that slot, is kept as a forest over the 32 old
// contents: merging R[r] into the builder links the root of R[r]'s tree under the
// root of the builder's tree (O(1) per token; the loops over all slots happen
// once, at the end of the tile).
//   root[d]    tree whose contents are in current slot d (ACT_DEAD: none)
//   parent[x]  ACT_DEAD for a root
//   delta[x]   end offset of x's content inside its slot = len0[x] + sum of delta along the path to the root
// fate[x] = slot the old content x ended up in (ACT_DEAD: dropped); `sor` is scratch.
#define ACT_FATES(root, parent, sor, fate)                                                    \
  {                                                                                           \
    for (uint32_t x_ = 0; x_ < nslots; ++x_) ACT_S(sor, x_) = (uint8_t)ACT_DEAD;              \
    for (uint32_t d_ = 0; d_ < nslots; ++d_) { const uint32_t r_ = ACT_S(root, d_); if (r_ != ACT_DEAD) ACT_S(sor, r_) = (uint8_t)d_; } \
    for (uint32_t x_ = 0; x_ < nslots; ++x_) {                                                \
      uint32_t r_ = (uint32_t)x_;                                                             \
      for (;;) { const uint32_t p_ = ACT_S(parent, r_); if (p_ == ACT_DEAD) break; r_ = p_; } \
      ACT_S(fate, x_) = ACT_S(sor, r_);                                                       \
    }                                                                                         \
  }

// A thread's row of a per-tile table is contiguous (32 B of fates, 128 B of
// lengths): whole-sector vector accesses.
// (only the first `nslots` slots are in use: slot h < nslots - registers = builder at height h, slot
// nslots-1-r = register r; the unused entries of a row are DEAD / 0 so that the warp kernels, lane = slot,
// need not know)
#define ACT_ST_ROW8(dst, arr)                                                                 \
  {                                                                                           \
    uint32_t w_[8];                                                                           \
    _Pragma("unroll")                                                                         \
    for (uint32_t k_ = 0; k_ < 8u; ++k_) {                                                    \
      uint32_t v_ = 0xFFFFFFFFu;                                                              \
      if (4u * k_ < nslots) {                                                                 \
        v_ = 0u;                                                                              \
        _Pragma("unroll")                                                                     \
        for (uint32_t j_ = 0; j_ < 4u; ++j_)                                                  \
          v_ |= (4u * k_ + j_ < nslots ? (uint32_t)ACT_S(arr, 4u * k_ + j_) : 0xFFu) << (8u * j_); \
      }                                                                                       \
      w_[k_] = v_;                                                                            \
    }                                                                                         \
    ((uint4 *)(dst))[0] = make_uint4(w_[0], w_[1], w_[2], w_[3]);                             \
    ((uint4 *)(dst))[1] = make_uint4(w_[4], w_[5], w_[6], w_[7]);                             \
  }
#define ACT_ST_ROW32(dst, arr)                                                                \
  {                                                                                           \
    _Pragma("unroll")                                                                         \
    for (uint32_t k_ = 0; k_ < 8u; ++k_) {                                                    \
      uint4 q_ = make_uint4(0u, 0u, 0u, 0u);                                                  \
      if (4u * k_ < nslots) {                                                                 \
        q_.x = ACT_S(arr, 4u * k_);                                                           \
        if (4u * k_ + 1u < nslots) q_.y = ACT_S(arr, 4u * k_ + 1u);                           \
        if (4u * k_ + 2u < nslots) q_.z = ACT_S(arr, 4u * k_ + 2u);                           \
        if (4u * k_ + 3u < nslots) q_.w = ACT_S(arr, 4u * k_ + 3u);                           \
      }                                                                                       \
      ((uint4 *)(dst))[k_] = q_;                                                              \
    }                                                                                         \
  }
#define ACT_LD_ROW32(src, arr)                                                                \
  {                                                                                           \
    for (uint32_t k_ = 0; 4u * k_ < nslots; ++k_) {                                           \
      const uint4 q_ = ((const uint4 *)(src))[k_];                                            \
      ACT_S(arr, 4u * k_) = q_.x;                                                             \
      if (4u * k_ + 1u < nslots) ACT_S(arr, 4u * k_ + 1u) = q_.y;                             \
      if (4u * k_ + 2u < nslots) ACT_S(arr, 4u * k_ + 2u) = q_.z;                             \
      if (4u * k_ + 3u < nslots) ACT_S(arr, 4u * k_ + 3u) = q_.w;                             \
    }                                                                                         \
  }

extern __shared__ __align__(16) uint32_t act_sm[];

// ---- P1: forward summary of a tile (fate of every slot's old content, bytes added per slot)
__global__ void __launch_bounds__(ACT_NT)
ka_fwd_summary(const uint8_t *__restrict__ in, size_t n, uint32_t tile, size_t ntiles,
               const int32_t *__restrict__ h0, uint32_t nslots, uint8_t *__restrict__ fate_out,
               uint32_t *__restrict__ add_out) {
  uint32_t *add = act_sm;                                        // dynamic shared memory: nslots * ACT_NT * (4 + 4 x 1) bytes
  uint8_t *root = (uint8_t *)(add + nslots * ACT_NT), *parent = root + nslots * ACT_NT, *sor = parent + nslots * ACT_NT,
          *fate = sor + nslots * ACT_NT;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntiles) return;
  const size_t lo = t * tile, hi = (lo + tile < n) ? lo + tile : n;
  for (uint32_t x = 0; x < nslots; ++x) { ACT_S(add, x) = 0; ACT_S(root, x) = (uint8_t)x; ACT_S(parent, x) = (uint8_t)ACT_DEAD; }
  uint32_t h = (uint32_t)h0[t];
  uint32_t cur = 0;                                 // add[h], kept in a register while h does not change
#define A_BYTE(i, v) { cur += 1u; }
#define A_PUSH(i) { ACT_S(add, h) = cur; ++h; cur = ACT_S(add, h); }
#define A_POP(i, r) { const uint32_t N = nslots - 1u - (r);                                \
    ACT_S(root, N) = ACT_S(root, h); ACT_S(root, h) = (uint8_t)ACT_DEAD;                      \
    ACT_S(add, N) = cur; ACT_S(add, h) = 0; --h; cur = ACT_S(add, h); }
#define A_WRITE(i, r) { const uint32_t N = nslots - 1u - (r);                              \
    const uint32_t rN = ACT_S(root, N);                                                       \
    if (rN != ACT_DEAD) {                                                                     \
      const uint32_t rh = ACT_S(root, h);                                                     \
      if (rh == ACT_DEAD) ACT_S(root, h) = (uint8_t)rN; else ACT_S(parent, rN) = (uint8_t)rh; \
      ACT_S(root, N) = (uint8_t)ACT_DEAD;                                                     \
    }                                                                                         \
    cur += ACT_S(add, N); ACT_S(add, N) = 0; }
  ACT_FWD_TOKENS(in, lo, hi, A_BYTE, A_PUSH, A_POP, A_WRITE)
#undef A_BYTE
#undef A_PUSH
#undef A_POP
#undef A_WRITE
  ACT_S(add, h) = cur;
  ACT_FATES(root, parent, sor, fate)
  ACT_ST_ROW8(fate_out + t * ACT_NSLOT, fate)
  ACT_ST_ROW32(add_out + t * ACT_NSLOT, add)
}

// The kernels below run one WARP per group of tiles, lane = slot.

// out[d] = base[d] + sum of v[m] over the slots m with f[m] == d  (lane = slot; `sm` = 32 words of this warp):
// one shared-memory atomic per lane instead of a 32-step shuffle loop
__device__ __forceinline__ uint32_t act_scatter_add(uint32_t *sm, uint32_t lane, uint32_t base, uint32_t f, uint32_t v) {
  sm[lane] = base;
  __syncwarp();
  if (f != ACT_DEAD) atomicAdd(&sm[f], v);
  __syncwarp();
  const uint32_t r = sm[lane];
  __syncwarp();
  return r;
}

// forward composition of the summaries of one group of tiles (earlier first)
__global__ void __launch_bounds__(128)
ka_group_compose(const uint8_t *__restrict__ fate, const uint32_t *__restrict__ add, size_t ntiles, size_t ngroups,
                 uint8_t *__restrict__ gfate, uint32_t *__restrict__ gadd) {
  __shared__ uint32_t sm[4][32];
  const size_t g = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  if (g >= ngroups) return;
  const size_t lo = g * ACT_GROUP, hi = (lo + ACT_GROUP < ntiles) ? lo + ACT_GROUP : ntiles;
  uint32_t f = fate[lo * ACT_NSLOT + lane], a = add[lo * ACT_NSLOT + lane];
  for (size_t t = lo + 1; t < hi; ++t) {
    const uint32_t f2 = fate[t * ACT_NSLOT + lane];
    const uint32_t acc = act_scatter_add(sm[threadIdx.x >> 5], lane, add[t * ACT_NSLOT + lane], f2, a);
    const uint32_t nf = __shfl_sync(0xFFFFFFFFu, f2, f & 31u);
    f = (f == ACT_DEAD) ? ACT_DEAD : nf;
    a = acc;
  }
  gfate[g * ACT_NSLOT + lane] = (uint8_t)f;
  gadd[g * ACT_NSLOT + lane] = a;
}

// one warp: slot lengths at the start of every group; gvec[ngroups] = at the end of the stream
__global__ void __launch_bounds__(32)
ka_group_scan(const uint8_t *__restrict__ gfate, const uint32_t *__restrict__ gadd, size_t ngroups,
              uint32_t *__restrict__ gvec) {
  __shared__ uint32_t sm[32];
  const uint32_t lane = threadIdx.x;
  uint32_t v = 0;
  uint32_t f = ngroups ? gfate[lane] : 0u, acc = ngroups ? gadd[lane] : 0u;
  for (size_t g = 0; g < ngroups; ++g) {
    gvec[g * ACT_NSLOT + lane] = v;
    // the next group's summary is loaded before this one's dependent arithmetic
    const uint32_t nf = (g + 1 < ngroups) ? gfate[(g + 1) * ACT_NSLOT + lane] : 0u;
    const uint32_t na = (g + 1 < ngroups) ? gadd[(g + 1) * ACT_NSLOT + lane] : 0u;
    v = act_scatter_add(sm, lane, acc, f, v);
    f = nf;
    acc = na;
  }
  gvec[ngroups * ACT_NSLOT + lane] = v;
}

// slot lengths at the start of every tile
__global__ void __launch_bounds__(128)
ka_tile_vectors(const uint8_t *__restrict__ fate, const uint32_t *__restrict__ add, size_t ntiles, size_t ngroups,
                const uint32_t *__restrict__ gvec, uint32_t *__restrict__ vec) {
  __shared__ uint32_t sm[4][32];
  const size_t g = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  if (g >= ngroups) return;
  const size_t lo = g * ACT_GROUP, hi = (lo + ACT_GROUP < ntiles) ? lo + ACT_GROUP : ntiles;
  uint32_t v = gvec[g * ACT_NSLOT + lane];
  for (size_t t = lo; t < hi; ++t) {
    vec[t * ACT_NSLOT + lane] = v;
    v = act_scatter_add(sm[threadIdx.x >> 5], lane, add[t * ACT_NSLOT + lane], fate[t * ACT_NSLOT + lane], v);
  }
}

// ---- P2: exact lengths; length of R[r] at every `write r`; backward summary of the tile
// (fate as in P1; tail = bytes behind a slot's old content inside the slot it ended up in)
__global__ void __launch_bounds__(ACT_NT)
ka_fwd_exact(const uint8_t *__restrict__ in, size_t n, uint32_t tile, size_t ntiles, const int32_t *__restrict__ h0,
             uint32_t nslots, const uint32_t *__restrict__ vec, uint32_t *__restrict__ wlen,
             uint8_t *__restrict__ fate_out, uint32_t *__restrict__ tail_out, ActCtl *ctl) {
  uint32_t *len = act_sm, *delta = len + nslots * ACT_NT;        // dynamic shared memory: nslots * ACT_NT * (8 + 4 x 1) bytes
  uint8_t *root = (uint8_t *)(delta + nslots * ACT_NT), *parent = root + nslots * ACT_NT, *sor = parent + nslots * ACT_NT,
          *fate = sor + nslots * ACT_NT;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntiles) return;
  const size_t lo = t * tile, hi = (lo + tile < n) ? lo + tile : n;
  ACT_LD_ROW32(vec + t * ACT_NSLOT, len)
  for (uint32_t x = 0; x < nslots; ++x) { ACT_S(delta, x) = 0; ACT_S(root, x) = (uint8_t)x; ACT_S(parent, x) = (uint8_t)ACT_DEAD; }
  uint32_t h = (uint32_t)h0[t];
  uint32_t cur = ACT_S(len, h);                     // len[h], kept in a register while h does not change
#define A_BYTE(i, v) { cur += 1u; }
#define A_PUSH(i) { ACT_S(len, h) = cur; ++h; cur = ACT_S(len, h); }
#define A_POP(i, r) { const uint32_t N = nslots - 1u - (r);                                \
    ACT_S(root, N) = ACT_S(root, h); ACT_S(root, h) = (uint8_t)ACT_DEAD;                      \
    ACT_S(len, N) = cur; ACT_S(len, h) = 0; --h; cur = ACT_S(len, h); }
#define A_WRITE(i, r) { const uint32_t N = nslots - 1u - (r);                              \
    const uint32_t ln_ = ACT_S(len, N);                                                       \
    wlen[(i) >> 1] = ln_;                                                                     \
    const uint32_t rN = ACT_S(root, N);                                                       \
    if (rN != ACT_DEAD) {                                                                     \
      /* R[r]'s contents now end `cur` bytes further into the builder */                     \
      const uint32_t rh = ACT_S(root, h);                                                     \
      if (rh == ACT_DEAD) { ACT_S(root, h) = (uint8_t)rN; ACT_S(delta, rN) += cur; }          \
      else { ACT_S(parent, rN) = (uint8_t)rh; ACT_S(delta, rN) += cur - ACT_S(delta, rh); }   \
      ACT_S(root, N) = (uint8_t)ACT_DEAD;                                                     \
    }                                                                                         \
    cur += ln_; ACT_S(len, N) = 0; }
  ACT_FWD_TOKENS(in, lo, hi, A_BYTE, A_PUSH, A_POP, A_WRITE)
#undef A_BYTE
#undef A_PUSH
#undef A_POP
#undef A_WRITE
  ACT_S(len, h) = cur;
  ACT_FATES(root, parent, sor, fate)
  // tail[x] = len[fate[x]] - (len0[x] + path sum of delta); it replaces delta[x] only after
  // every path has been summed (paths run through other slots' deltas)
  uint32_t tl[ACT_NSLOT];
  const uint32_t *len0 = vec + t * ACT_NSLOT;
#pragma unroll 1
  for (uint32_t x = 0; x < nslots; ++x) {
    const uint32_t d = ACT_S(fate, x);
    uint32_t endoff = len0[x], r = (uint32_t)x;
    for (;;) {
      endoff += ACT_S(delta, r);
      const uint32_t pr = ACT_S(parent, r);
      if (pr == ACT_DEAD) break;
      r = pr;
    }
    tl[x] = (d == ACT_DEAD) ? 0u : ACT_S(len, d) - endoff;
  }
  for (uint32_t x = 0; x < nslots; ++x) ACT_S(delta, x) = tl[x];
  ACT_ST_ROW8(fate_out + t * ACT_NSLOT, fate)
  ACT_ST_ROW32(tail_out + t * ACT_NSLOT, delta)
  if (t == ntiles - 1) ctl->total = ACT_S(len, 0);
}

// ---- P3: backward composition (earlier tile first, then the later one)
__global__ void __launch_bounds__(128)
ka_group_bcompose(const uint8_t *__restrict__ fate, const uint32_t *__restrict__ tail, size_t ntiles, size_t ngroups,
                  uint8_t *__restrict__ gfate, uint32_t *__restrict__ gtail) {
  const size_t g = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  if (g >= ngroups) return;
  const size_t lo = g * ACT_GROUP, hi = (lo + ACT_GROUP < ntiles) ? lo + ACT_GROUP : ntiles;
  uint32_t f = fate[lo * ACT_NSLOT + lane], o = tail[lo * ACT_NSLOT + lane];
  for (size_t t = lo + 1; t < hi; ++t) {
    const uint32_t f2 = fate[t * ACT_NSLOT + lane], o2 = tail[t * ACT_NSLOT + lane];
    const uint32_t d2 = __shfl_sync(0xFFFFFFFFu, f2, f & 31u), od = __shfl_sync(0xFFFFFFFFu, o2, f & 31u);
    if (f != ACT_DEAD) {
      if (d2 == ACT_DEAD) { f = ACT_DEAD; o = 0; }
      else { f = d2; o += od; }
    }
  }
  gfate[g * ACT_NSLOT + lane] = (uint8_t)f;
  gtail[g * ACT_NSLOT + lane] = o;
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
  const size_t g = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  if (g >= ngroups) return;
  const size_t lo = g * ACT_GROUP, hi = (lo + ACT_GROUP < ntiles) ? lo + ACT_GROUP : ntiles;
  uint32_t e = gvec[g * ACT_NSLOT + lane];
  for (size_t t = hi; t-- > lo;) {
    vec[t * ACT_NSLOT + lane] = e;
    const uint32_t f = fate[t * ACT_NSLOT + lane], o = tail[t * ACT_NSLOT + lane];
    const uint32_t ef = __shfl_sync(0xFFFFFFFFu, e, f & 31u);
    e = (f == ACT_DEAD || ef == ACT_NOPOS) ? ACT_NOPOS : ef - o;
  }
}

// ---- P4: backward walk of the tile; every surviving byte is stored at its final position
__global__ void __launch_bounds__(ACT_NT)
ka_write(const uint8_t *__restrict__ in, size_t n, uint32_t tile, size_t ntiles, const int32_t *__restrict__ h0,
         uint32_t nslots, const uint32_t *__restrict__ vec, const uint32_t *__restrict__ wlen,
         uint8_t *__restrict__ out) {
  uint32_t *e = act_sm;                                          // dynamic shared memory: nslots * ACT_NT * 4 bytes
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntiles) return;
  const size_t lo = t * tile, hi = (lo + tile < n) ? lo + tile : n;
  if (hi <= lo) return;
  ACT_LD_ROW32(vec + t * ACT_NSLOT, e)
  uint32_t h = (uint32_t)h0[t + 1];
  uint32_t eh = ACT_S(e, h);                          // e[h], kept in a register while h does not change
  // Consecutive bytes of a run go to descending addresses: they are collected in
  // `acc` (nacc bytes, the lowest at out[eh]) and leave as one aligned word when a
  // word is complete, as single bytes when the run ends inside a word.
  uint32_t acc = 0, nacc = 0;
#ifdef ACT_BYTE_STORES          // plain byte stores, kept for comparison: 30.7 vs 33.4 GiB/s on swap_fields
#define ACT_PUT(v) { out[--eh] = (uint8_t)(v); }
#define ACT_FLUSH() {}
#else
#define ACT_PUT(v)                                                                            \
  {                                                                                           \
    --eh;                                                                                     \
    acc = (acc << 8) | (v);                                                                   \
    ++nacc;                                                                                   \
    if ((eh & 3u) == 0u) {                                                                    \
      if (nacc == 4u) *(uint32_t *)(out + eh) = acc;                                          \
      else for (uint32_t j_ = 0; j_ < nacc; ++j_) out[eh + j_] = (uint8_t)(acc >> (8u * j_)); \
      acc = 0; nacc = 0;                                                                      \
    }                                                                                         \
  }
#define ACT_FLUSH()                                                                           \
  {                                                                                           \
    for (uint32_t j_ = 0; j_ < nacc; ++j_) out[eh + j_] = (uint8_t)(acc >> (8u * j_));        \
    acc = 0; nacc = 0;                                                                        \
  }
#endif
  size_t c = (hi - 1) & ~(size_t)3;
  uint32_t cnt = (uint32_t)(hi - c);
  uint32_t w = act_ld4(in, c, cnt), pw = 0;
  bool skip = false;                                  // the byte is the ESC lead of the operand just handled
#pragma unroll 1
  for (;;) {
    // the word before this one supplies the byte in front of this word's first byte
    uint32_t before = 0u;
    if (c > lo) { pw = __ldg((const uint32_t *)(in + c - 4)); before = pw >> 24; }
    else if (c > 0) before = in[c - 1];
#pragma unroll
    for (int k = 3; k >= 0; --k) {
      if ((uint32_t)k < cnt) {
        const uint32_t b = (w >> (8 * k)) & 0xFFu;
        const uint32_t pb = k > 0 ? (w >> (8 * (k - 1))) & 0xFFu : before;
        if (skip) {
          skip = false;
        } else if (pb == ACT_ESC) {                   // an operand
          skip = true;
          if (b == 0u) {
            if (eh != ACT_NOPOS) ACT_PUT(ACT_ESC)
          } else {
            ACT_FLUSH()
            if (b == 1u) {                            // undo push
              ACT_S(e, h) = eh; --h; eh = ACT_S(e, h);
            } else if (b & 1u) {                      // undo write r
              const uint32_t N = nslots - 1u - ((b - 3u) >> 1);
              ACT_S(e, N) = eh;
              if (eh != ACT_NOPOS) eh -= wlen[(c + (size_t)k) >> 1];
            } else {                                  // undo pop r
              const uint32_t N = nslots - 1u - ((b - 2u) >> 1);
              ACT_S(e, h) = eh; ++h;
              eh = ACT_S(e, N);
              ACT_S(e, N) = ACT_NOPOS;
            }
          }
        } else if (b != ACT_ESC) {
          if (eh != ACT_NOPOS) ACT_PUT(b)
        }
      }
    }
    if (c <= lo) break;
    c -= 4;
    cnt = 4u;
    w = pw;
  }
  ACT_FLUSH()
#undef ACT_PUT
#undef ACT_FLUSH
}
