/*
 * ORACLE / TEST INFRASTRUCTURE -- not part of the product path.
 *
 * Plain-C restatement of the reference's sequential hot path: the per-state
 * test chain of a generated `matchN()` (src/KMC/Program/Backends/C.hs:72-83,
 * block shape src/KMC/SSTCompiler.hs:132-156) interpreted from a serialised
 * SST, over a restatement of the runtime primitives of crt/crt.c:
 *   growable register buffers  init_buffer/reset/append/appendarray/concat
 *                              (crt/crt.c:161-259; concat copies, never moves)
 *   output stream              outputconst/outputarray/output with a flush
 *                              every 16 KiB (crt/crt.c:107-159,217-283,334-354)
 *   reject                     "Match error at input symbol <count>" + only the
 *                              whole flushes reach the stream (C.hs:79-81)
 * Validated against oracle/_ref binaries (emitted C + verbatim crt.c) and the
 * reference's golden vectors in tests/test_oracle.py.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 *
 * Serialised SST ("KSST", little-endian u32 words; oracle/sstbin.py writes it):
 *   magic, nstates, nvars, init
 *   per state: final(0|1), natoms, atoms..., ntrans,
 *              per transition: pred[8] (256-bit set), dest, nassign,
 *                              per assignment (execution order): var, natoms, atoms...
 *   atom: 0 var | 1 len bytes(padded to words) | 2 (current input byte)
 *         | 3 table[256] (64 words): table[current input byte] -- AppendTblI of the
 *           oracle / action programs of the default --act=true mode (C.hs:232-251)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define OUTBUFFER_SIZE (16 * 1024)
#define INITIAL_BUFFER_SIZE (4096 * 8)

typedef struct { uint8_t *data; size_t size, len; } regbuf;

typedef struct { const uint32_t *atoms; uint32_t natoms, var; } assign_t;
typedef struct { uint32_t dest, nassign; assign_t *as; } trans_t;
typedef struct { int final; const uint32_t *fatoms; uint32_t nfatoms; trans_t *tr; uint32_t ntr; int32_t disp[256]; } state_t;

static const uint32_t *skip_atoms(const uint32_t *p, uint32_t n) {
  for (uint32_t i = 0; i < n; ++i) {
    uint32_t t = *p++;
    if (t == 0) p++;
    else if (t == 1) { uint32_t l = *p++; p += (l + 3) / 4; }
    else if (t == 3) p += 64;
  }
  return p;
}

static void buf_reserve(regbuf *b, size_t total) {
  /* appendarray's doubling rule (crt/crt.c:229-244) */
  if (total >= b->size - 1) {
    size_t ns = b->size;
    while (total >= ns - 1) ns <<= 1;
    uint8_t *d = (uint8_t *)calloc(ns, 1);
    memcpy(d, b->data, b->len);
    free(b->data);
    b->data = d;
    b->size = ns;
  }
}

typedef struct { uint8_t *out; size_t cap, flushed, pend; int overflow; uint8_t win[OUTBUFFER_SIZE]; } outstream;

static inline void out_byte(outstream *o, uint8_t w) {
  /* outputconst -> buf_writeconst -> buf_flush at 16 KiB (crt/crt.c:140-159,217-227) */
  o->win[o->pend++] = w;
  if (o->pend == OUTBUFFER_SIZE) {
    if (o->flushed + o->pend <= o->cap) memcpy(o->out + o->flushed, o->win, o->pend);
    else o->overflow = 1;
    o->flushed += o->pend;
    o->pend = 0;
  }
}

static void run_atoms(const uint32_t *p, uint32_t n, uint32_t self, int to_stream, regbuf *regs, outstream *o,
                      uint8_t sym, int skip_self_head) {
  regbuf *dst = to_stream ? NULL : &regs[self];
  for (uint32_t i = 0; i < n; ++i) {
    uint32_t t = *p++;
    if (t == 0) {
      uint32_t v = *p++;
      if (i == 0 && skip_self_head && v == self) continue;
      if (to_stream) { for (size_t k = 0; k < regs[v].len; ++k) out_byte(o, regs[v].data[k]); }   /* output() */
      else { buf_reserve(dst, dst->len + regs[v].len); memcpy(dst->data + dst->len, regs[v].data, regs[v].len); dst->len += regs[v].len; } /* concat() */
    } else if (t == 1) {
      uint32_t l = *p++;
      const uint8_t *c = (const uint8_t *)p;
      p += (l + 3) / 4;
      if (to_stream) { for (uint32_t k = 0; k < l; ++k) out_byte(o, c[k]); }                      /* outputarray() */
      else { buf_reserve(dst, dst->len + l); memcpy(dst->data + dst->len, c, l); dst->len += l; }  /* appendarray() */
    } else {
      uint8_t w = sym;
      if (t == 3) { w = ((const uint8_t *)p)[sym]; p += 64; }                                      /* tblP[t][next[0]] */
      if (to_stream) out_byte(o, w);                                                               /* outputconst(next[0],8) */
      else { buf_reserve(dst, dst->len + 1); dst->data[dst->len++] = w; }                          /* append() */
    }
  }
}

/* returns 0 accept, 1 reject, <0 malformed program / output overflow (-3: *outlen = needed) */
static int oracle_run_impl(const void *sst, size_t sstlen, const uint8_t *in, size_t n, uint8_t *out, size_t cap,
                           size_t *outlen, size_t *count, int keep_pending);

int kex_oracle_run(const void *sst, size_t sstlen, const uint8_t *in, size_t n, uint8_t *out, size_t cap,
                   size_t *outlen, size_t *count) {
  return oracle_run_impl(sst, sstlen, in, n, out, cap, outlen, count, 0);
}

/* The transducer phase of a stage with register actions (frontend/actions.py): its output is not
 * the program's stdout but the action stream handed to the action interpreter inside the same
 * stage, so on a reject everything produced up to the failing symbol is handed on; the 16 KiB
 * flush rule applies once, to the interpreted output (oracle/sstbin.py). */
int kex_oracle_run_stream(const void *sst, size_t sstlen, const uint8_t *in, size_t n, uint8_t *out, size_t cap,
                          size_t *outlen, size_t *count) {
  return oracle_run_impl(sst, sstlen, in, n, out, cap, outlen, count, 1);
}

static int oracle_run_impl(const void *sst, size_t sstlen, const uint8_t *in, size_t n, uint8_t *out, size_t cap,
                           size_t *outlen, size_t *count, int keep_pending) {
  const uint32_t *p = (const uint32_t *)sst;
  if (sstlen < 16 || p[0] != 0x5453534Bu) return -1;
  const uint32_t nstates = p[1], nvars = p[2], init = p[3];
  p += 4;
  state_t *S = (state_t *)calloc(nstates, sizeof(state_t));
  for (uint32_t q = 0; q < nstates; ++q) {
    S[q].final = (int)*p++;
    S[q].nfatoms = *p++;
    S[q].fatoms = p;
    p = skip_atoms(p, S[q].nfatoms);
    S[q].ntr = *p++;
    S[q].tr = (trans_t *)calloc(S[q].ntr ? S[q].ntr : 1, sizeof(trans_t));
    for (int b = 0; b < 256; ++b) S[q].disp[b] = -1;
    for (uint32_t t = 0; t < S[q].ntr; ++t) {
      const uint32_t *pred = p;
      p += 8;
      /* first matching test wins, as in the generated if-chain (C.hs:340-368) */
      for (int b = 0; b < 256; ++b)
        if (((pred[b >> 5] >> (b & 31)) & 1u) && S[q].disp[b] < 0) S[q].disp[b] = (int32_t)t;
      S[q].tr[t].dest = *p++;
      S[q].tr[t].nassign = *p++;
      S[q].tr[t].as = (assign_t *)calloc(S[q].tr[t].nassign ? S[q].tr[t].nassign : 1, sizeof(assign_t));
      for (uint32_t a = 0; a < S[q].tr[t].nassign; ++a) {
        S[q].tr[t].as[a].var = *p++;
        S[q].tr[t].as[a].natoms = *p++;
        S[q].tr[t].as[a].atoms = p;
        p = skip_atoms(p, S[q].tr[t].as[a].natoms);
      }
    }
  }
  regbuf *regs = (regbuf *)calloc(nvars ? nvars : 1, sizeof(regbuf));
  for (uint32_t v = 1; v < nvars; ++v) {   /* init_buffer for every non-stream buffer (C.hs:470-484) */
    regs[v].data = (uint8_t *)malloc(INITIAL_BUFFER_SIZE);
    regs[v].size = INITIAL_BUFFER_SIZE;
    regs[v].len = 0;
  }
  outstream *o = (outstream *)calloc(1, sizeof(outstream));
  o->out = out;
  o->cap = cap;
  uint32_t q = init;
  size_t i = 0;
  int status = 0;
  for (;;) {
    if (i == n) {   /* readnext(1,1) fails: end-of-input branch (SSTCompiler.hs:147-154) */
      if (!S[q].final) { status = 1; break; }
      run_atoms(S[q].fatoms, S[q].nfatoms, 0, 1, regs, o, 0, 0);
      break;
    }
    const uint8_t b = in[i];
    const int32_t t = S[q].disp[b];
    if (t < 0) { status = 1; break; }   /* goto failN */
    const trans_t *tr = &S[q].tr[t];
    for (uint32_t a = 0; a < tr->nassign; ++a) {
      const assign_t *as = &tr->as[a];
      /* compileAssignment: reset unless the update starts with the register itself (SSTCompiler.hs:67-80) */
      int self_head = as->natoms && as->atoms[0] == 0 && as->atoms[1] == as->var;
      if (as->var == 0) run_atoms(as->atoms, as->natoms, 0, 1, regs, o, b, 1);
      else {
        if (!self_head) regs[as->var].len = 0;
        run_atoms(as->atoms, as->natoms, as->var, 0, regs, o, b, self_head);
      }
    }
    ++i;          /* consume(1) */
    q = tr->dest;
  }
  *count = i;
  if (status == 0 || keep_pending) {   /* flush_outbuf (crt/crt.c:334-346) */
    if (o->flushed + o->pend <= o->cap) memcpy(o->out + o->flushed, o->win, o->pend);
    else o->overflow = 1;
    o->flushed += o->pend;
  }
  *outlen = o->flushed;
  const int ovf = o->overflow;
  for (uint32_t v = 1; v < nvars; ++v) free(regs[v].data);
  free(regs);
  for (uint32_t s = 0; s < nstates; ++s) {
    for (uint32_t t = 0; t < S[s].ntr; ++t) free(S[s].tr[t].as);
    free(S[s].tr);
  }
  free(S);
  free(o);
  return ovf ? -3 : status;
}

/*
 * Action interpreter over an action stream (ORACLE, as above): a stack of
 * output builders plus a register bank -- the action semantics of
 * src/KMC/Kleenex/Actions.hs:14-58, i.e. what `interp` makes of the symbols in
 * src/KMC/SymbolicSST/ActionSST.hs:83-104.  Stream encoding: see
 * kleenexlang_b200/frontend/actions.py (ESC = 0xFF; ESC 0 literal, ESC 1 push,
 * ESC 2+2r pop r, ESC 3+2r write r).  The result is the bottom builder.
 * Returns 0, -3 (output does not fit; *outlen = needed) or -1 (malformed).
 */
typedef struct { uint8_t *d; size_t len, cap; } abuf;
static void abuf_put(abuf *b, const uint8_t *s, size_t n) {
  if (b->len + n > b->cap) {
    size_t c = b->cap ? b->cap : 64;
    while (c < b->len + n) c *= 2;
    b->d = (uint8_t *)realloc(b->d, c);
    b->cap = c;
  }
  if (n) memcpy(b->d + b->len, s, n);
  b->len += n;
}
int kex_oracle_act(const uint8_t *in, size_t n, uint32_t nregs, uint8_t *out, size_t cap, size_t *outlen) {
  size_t depth = 0, scap = 8;
  abuf *stack = (abuf *)calloc(scap, sizeof(abuf));
  abuf *regs = (abuf *)calloc(nregs ? nregs : 1, sizeof(abuf));
  int rc = 0;
  for (size_t i = 0; i < n && rc == 0; ++i) {
    const uint8_t b = in[i];
    if (b != 0xFF) { abuf_put(&stack[depth], &b, 1); continue; }
    if (i + 1 >= n) break;                      /* truncated stream: lone ESC */
    const uint32_t c = in[++i];
    if (c == 0) { abuf_put(&stack[depth], &b, 1); }
    else if (c == 1) {
      if (++depth == scap) {
        stack = (abuf *)realloc(stack, 2 * scap * sizeof(abuf));
        memset(stack + scap, 0, scap * sizeof(abuf));
        scap *= 2;
      }
      stack[depth].len = 0;
    } else {
      const uint32_t r = (c & 1) ? (c - 3) / 2 : (c - 2) / 2;
      if (r >= nregs) { rc = -1; break; }
      if (c & 1) {                              /* write r: append and clear */
        abuf_put(&stack[depth], regs[r].d, regs[r].len);
        regs[r].len = 0;
      } else {                                  /* pop r: the top builder becomes r */
        if (depth == 0) { rc = -1; break; }
        abuf t = regs[r];
        regs[r] = stack[depth];
        stack[depth] = t;
        stack[depth].len = 0;
        --depth;
      }
    }
  }
  if (rc == 0) {
    *outlen = stack[0].len;
    if (stack[0].len > cap) rc = -3;
    else if (stack[0].len) memcpy(out, stack[0].d, stack[0].len);
  }
  for (size_t k = 0; k < scap; ++k) free(stack[k].d);
  for (uint32_t r = 0; r < nregs; ++r) free(regs[r].d);
  free(stack);
  free(regs);
  return rc;
}
