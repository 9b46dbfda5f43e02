/* kexrun.c -- native launcher for a program compiled by `kexc compile`: the
 * process contract of the reference's compiled binary (crt/crt.c:372-467:
 * `main`, option parsing :380-399, usage :326-332, timing :457-464; reject
 * message src/KMC/Program/Backends/C.hs:79-81) served through the C ABI of
 * libkexcuda.so (include/kexcuda.h).  No Python in the artefact.
 *
 *   ./bin < in > out        exit 0, or exit 1 + "Match error at input symbol N!"
 *   ./bin -i                compilation info, exit 2
 *   ./bin -h                usage, exit 1
 *   ./bin -t                also "time (ms): N" on stderr
 *   ./bin -p N | --phase N  runs only phase N
 *
 * `kexc compile --out bin` copies this executable to `bin` and writes
 * `bin.kexprog` (the program blob), `bin.kexprog.desc` (text of -i) and
 * `bin.kexprog.lib` (path of libkexcuda.so) next to it.  Also usable as
 * `kexrun --blob prog.kexprog [options]`.
 *
 * Inputs up to KEX_STREAM_BLOCK_MIB (default 256) MiB are evaluated by one
 * kex_run_host call; larger ones block by block with bounded memory
 * (kex_stream_*), writing -- like the reference runtime, crt/crt.c:107-159,
 * 217-227 -- only whole 16 KiB flushes until the run accepts.
 *
 * build: cc -O2 -Iinclude tools/kexrun.c -ldl -o kleenexlang_b200/kexrun
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <limits.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <unistd.h>

#include "kexcuda.h"

#define RETC_PRINT_USAGE 1 /* crt/crt.c:14 */
#define RETC_PRINT_INFO 2  /* crt/crt.c:15 */
#define FLUSH_UNIT 16384u  /* crt/crt.c:350 OUTBUFFER_SIZE */

static struct {
  int (*load)(const void *, size_t, int, kex_program **);
  void (*free_)(kex_program *);
  int (*run_host)(kex_program *, const uint8_t *, size_t, uint8_t *, size_t, size_t *, int *, size_t *);
  int (*select_phase)(kex_program *, uint32_t);
  size_t (*out_bound)(const kex_program *, size_t);
  int (*stream_begin)(kex_program *);
  int (*stream_feed)(kex_program *, const uint8_t *, size_t, uint8_t *, size_t, size_t *);
  int (*stream_end)(kex_program *, uint8_t *, size_t, size_t *, int *, size_t *);
  const char *(*strerror_)(int);
  const char *(*last_cuda_error)(const kex_program *);
} K;

static void die(const char *msg, const char *arg) {
  fprintf(stderr, "kexrun: %s%s%s\n", msg, arg ? ": " : "", arg ? arg : "");
  exit(1);
}

static uint8_t *read_file(const char *path, size_t *len, int must) {
  FILE *f = fopen(path, "rb");
  if (!f) {
    if (must) die("cannot open", path);
    return NULL;
  }
  size_t cap = 1 << 16, n = 0;
  uint8_t *buf = malloc(cap);
  for (;;) {
    if (n == cap) buf = realloc(buf, cap *= 2);
    size_t r = fread(buf + n, 1, cap - n, f);
    if (!r) break;
    n += r;
  }
  fclose(f);
  *len = n;
  return buf;
}

/* fills buf with up to cap bytes of stdin; < cap only at end of input */
static size_t read_full(uint8_t *buf, size_t cap) {
  size_t n = 0;
  while (n < cap) {
    size_t r = fread(buf + n, 1, cap - n, stdin);
    if (!r) break;
    n += r;
  }
  return n;
}

static void load_library(const char *blob_path) {
  char path[PATH_MAX + 64];
  const char *cand[3] = {getenv("KEX_LIB"), NULL, "libkexcuda.so"};
  snprintf(path, sizeof path, "%s.lib", blob_path);
  size_t ll = 0;
  char *rec = (char *)read_file(path, &ll, 0);
  if (rec) {
    while (ll && (rec[ll - 1] == '\n' || rec[ll - 1] == '\r')) ll--;
    rec = realloc(rec, ll + 1);
    rec[ll] = 0;
    cand[1] = rec;
  }
  void *h = NULL;
  for (int i = 0; i < 3 && !h; ++i)
    if (cand[i] && *cand[i]) h = dlopen(cand[i], RTLD_NOW | RTLD_GLOBAL);
  if (!h) die("cannot load libkexcuda.so (set KEX_LIB; there is no CPU fallback)", dlerror());
#define SYM(field, name)                        \
  do {                                          \
    *(void **)(&K.field) = dlsym(h, name);      \
    if (!K.field) die("missing symbol", name);  \
  } while (0)
  SYM(load, "kex_load");
  SYM(free_, "kex_free");
  SYM(run_host, "kex_run_host");
  SYM(select_phase, "kex_select_phase");
  SYM(out_bound, "kex_out_bound");
  SYM(stream_begin, "kex_stream_begin");
  SYM(stream_feed, "kex_stream_feed");
  SYM(stream_end, "kex_stream_end");
  SYM(strerror_, "kex_strerror");
  SYM(last_cuda_error, "kex_last_cuda_error");
#undef SYM
}

static void fail_rc(kex_program *p, int rc) {
  fprintf(stderr, "kexrun: %s", K.strerror_(rc));
  if (rc == KEX_ERR_CUDA) fprintf(stderr, ": %s", K.last_cuda_error(p));
  fputc('\n', stderr);
  exit(1);
}

static void usage(const char *name) { /* crt/crt.c:326-332 */
  printf("Normal usage: %s < infile > outfile\n", name);
  printf("- \"%s -i\": Print compilation info\n", name);
  printf("- \"%s -t\": Runs normally, but prints timing to stderr\n", name);
  printf("- \"%s -p N\": Runs only phase N\n", name);
}

/* output held back so that a rejecting run leaves only whole 16 KiB flushes on stdout */
static uint8_t *held;
static size_t held_n, held_cap;
static void push(const uint8_t *d, size_t n, int everything) {
  if (held_n + n > held_cap) held = realloc(held, held_cap = (held_n + n) * 2 + FLUSH_UNIT);
  memcpy(held + held_n, d, n);
  held_n += n;
  size_t keep = everything ? 0 : held_n % FLUSH_UNIT;
  if (held_n > keep) {
    fwrite(held, 1, held_n - keep, stdout);
    memmove(held, held + held_n - keep, keep);
    held_n = keep;
  }
}

int main(int argc, char **argv) {
  char exe[PATH_MAX], blob_path[PATH_MAX + 16];
  const char *blob_arg = NULL;
  int do_timing = 0, first_opt = 1;
  long phase = 0;
  if (argc >= 3 && !strcmp(argv[1], "--blob")) {
    blob_arg = argv[2];
    first_opt = 3;
  }
  if (blob_arg) {
    snprintf(blob_path, sizeof blob_path, "%s", blob_arg);
  } else {
    ssize_t l = readlink("/proc/self/exe", exe, sizeof exe - 1);
    if (l <= 0) die("cannot resolve /proc/self/exe", NULL);
    exe[l] = 0;
    snprintf(blob_path, sizeof blob_path, "%s.kexprog", exe);
  }
  for (int i = first_opt; i < argc; ++i) { /* crt/crt.c:380-399 */
    const char *a = argv[i];
    if (!strcmp(a, "-i")) {
      char desc[PATH_MAX + 32];
      size_t dl = 0;
      snprintf(desc, sizeof desc, "%s.desc", blob_path);
      uint8_t *d = read_file(desc, &dl, 0);
      if (d) fwrite(d, 1, dl, stdout);
      else printf("Compiler info:\n  back end: CUDA (libkexcuda.so, sm_100a)\n  program: %s\n", blob_path);
      return RETC_PRINT_INFO;
    } else if (!strcmp(a, "-t")) {
      do_timing = 1;
    } else if ((!strcmp(a, "-p") || !strcmp(a, "--phase")) && i + 1 < argc) {
      phase = atol(argv[++i]);
    } else if (!strncmp(a, "--phase=", 8)) {
      phase = atol(a + 8);
    } else if (!strncmp(a, "-p", 2) && a[2]) {
      phase = atol(a + 2);
    } else { /* -h and anything unknown */
      usage(argv[0]);
      return RETC_PRINT_USAGE;
    }
  }
  struct timeval t0, t1;
  gettimeofday(&t0, NULL);
  size_t blob_len = 0;
  uint8_t *blob = read_file(blob_path, &blob_len, 1);
  load_library(blob_path);
  kex_program *h = NULL;
  int rc = K.load(blob, blob_len, getenv("KEX_DEVICE") ? atoi(getenv("KEX_DEVICE")) : 0, &h);
  if (rc) fail_rc(NULL, rc);
  if (phase) {
    if (K.select_phase(h, (uint32_t)phase)) { /* C.hs:59-69 */
      fprintf(stderr, "Invalid phase: %ld given\n", phase);
      return 1;
    }
  }
  size_t block = 256u << 20;
  if (getenv("KEX_STREAM_BLOCK_MIB") && atol(getenv("KEX_STREAM_BLOCK_MIB")) > 0)
    block = (size_t)atol(getenv("KEX_STREAM_BLOCK_MIB")) << 20;
  uint8_t *in = malloc(block + 1);
  if (!in) die("out of memory", NULL);
  size_t n = read_full(in, block + 1);
  int status = KEX_ACCEPT;
  size_t count = 0, out_len = 0;
  int streaming = n > block && !phase && !getenv("KEX_NO_STREAM") && K.stream_begin(h) == KEX_OK;
  if (!streaming) {
    /* the whole input at once (also: multi-phase programs and tables the streaming entry points do not serve) */
    size_t cap = block + 1;
    while (n == cap) { /* more input may be waiting */
      in = realloc(in, cap *= 2);
      if (!in) die("out of memory", NULL);
      n += read_full(in + n, cap - n);
    }
    size_t ocap = K.out_bound(h, n) + 4096;
    uint8_t *out = malloc(ocap);
    if (!out) die("out of memory", NULL);
    rc = K.run_host(h, in, n, out, ocap, &out_len, &status, &count);
    if (rc == KEX_ERR_OUT_CAP) {
      ocap = out_len + 4096;
      out = realloc(out, ocap);
      rc = K.run_host(h, in, n, out, ocap, &out_len, &status, &count);
    }
    if (rc) fail_rc(h, rc);
    fwrite(out, 1, out_len, stdout);
  } else {
    size_t ocap = 3 * block + 4096, wrote = 0;
    uint8_t *out = malloc(ocap);
    /* until the first output byte exists the fed blocks are kept: if the library then reports that
       the input cannot be streamed (a register stays live across blocks, KEX_ERR_UNSUPPORTED), the
       rest of stdin is read and everything is evaluated at once, as the reference binary would */
    uint8_t *spool = NULL;
    size_t spool_n = 0, spool_cap = 0;
    size_t have = n; /* block + 1 bytes: feed `block`, keep the extra byte as the head of the next block */
    for (;;) {
      size_t feed = have > block ? block : have;
      for (;;) {
        rc = K.stream_feed(h, in, feed, out, ocap, &out_len);
        if (rc != KEX_ERR_OUT_CAP) break;
        out = realloc(out, ocap = out_len + 4096);
      }
      if (rc == KEX_ERR_UNSUPPORTED) {
        if (wrote) die("this input cannot be evaluated block by block (a register stays live across blocks); "
                       "rerun with KEX_NO_STREAM=1", NULL);
        size_t cap = spool_n + have + block, m = spool_n + have;
        uint8_t *all = malloc(cap);
        if (!all) die("out of memory", NULL);
        if (spool_n) memcpy(all, spool, spool_n);
        memcpy(all + spool_n, in, have);
        for (;;) {
          if (m == cap) all = realloc(all, cap *= 2);
          size_t r = read_full(all + m, cap - m);
          if (!r) break;
          m += r;
        }
        K.free_(h);
        if ((rc = K.load(blob, blob_len, 0, &h))) fail_rc(NULL, rc);
        size_t oc = K.out_bound(h, m) + 4096;
        uint8_t *o2 = malloc(oc);
        if (!o2) die("out of memory", NULL);
        if ((rc = K.run_host(h, all, m, o2, oc, &out_len, &status, &count))) fail_rc(h, rc);
        fwrite(o2, 1, out_len, stdout);
        goto done;
      }
      if (rc) fail_rc(h, rc);
      if (!wrote && !out_len) {
        if (spool_n + feed > spool_cap) spool = realloc(spool, spool_cap = (spool_n + feed) * 2);
        memcpy(spool + spool_n, in, feed);
        spool_n += feed;
      } else if (spool) {
        free(spool);
        spool = NULL;
        spool_n = spool_cap = 0;
      }
      wrote += out_len;
      push(out, out_len, 0);
      size_t head = have > feed ? 1 : 0;
      if (head) in[0] = in[feed];
      size_t got = read_full(in + head, block + 1 - head);
      have = head + got;
      if (!have) break;
    }
    for (;;) {
      rc = K.stream_end(h, out, ocap, &out_len, &status, &count);
      if (rc != KEX_ERR_OUT_CAP) break;
      out = realloc(out, ocap = out_len + 4096);
    }
    if (rc) fail_rc(h, rc);
    push(out, out_len, status == KEX_ACCEPT);
  }
done:
  fflush(stdout);
  if (status != KEX_ACCEPT) {
    fprintf(stderr, "Match error at input symbol %zu!\n", count); /* C.hs:79-81 */
    return 1;
  }
  if (do_timing) { /* crt/crt.c:457-464 */
    gettimeofday(&t1, NULL);
    long ms = (t1.tv_sec - t0.tv_sec) * 1000 + (t1.tv_usec - t0.tv_usec) / 1000;
    fprintf(stderr, "time (ms): %ld\n", ms);
  }
  K.free_(h);
  return 0;
}
