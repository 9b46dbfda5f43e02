/*
 * kexcuda.h -- C ABI of libkexcuda.so, the B200 (sm_100a) execution back end
 * for compiled Kleenex streaming string transducers.
 *
 * The reference has no FFI for this path: the seam is the generated-code <->
 * runtime ABI of crt/crt.c:103-105 (`void printCompilationInfo(); void init();
 * void match(int phase);`) plus the process contract of the compiled binary
 * (crt/crt.c:372-467: stdin -> stdout, exit 0 on accept, exit 1 and
 * "Match error at input symbol <count>!" on reject).  The entry points below
 * are what a `KMC.Program.Backends.CUDA` module (the sibling of
 * src/KMC/Program/Backends/C.hs:529-540 `compileProgram`) would bind instead
 * of piping C text to `cc`: it serialises the pipeline's tables into a
 * `kexprog` blob (kleenexlang_b200/kexprog.py documents the layout) and the
 * launcher hands the blob and the input bytes to this library.
 *
 * A stage with register actions (`reg@t`, `!reg`: what compileOracleAction,
 * src/KMC/Frontend/Commands.hs:204-244, compiles to an oracle/action pair)
 * occupies two phases of a blob as well: a transducer phase whose output is an
 * action stream and an action-interpreter phase (the register effects of
 * src/KMC/SymbolicSST/ActionSST.hs:83-104; kleenexlang_b200/csrc/kex_act.cuh).
 * kex_run_host / kex_run_device / kex_select_phase / kex_info / kex_out_bound
 * handle such blobs; the sharded and block-streaming entry points answer
 * KEX_ERR_UNSUPPORTED (they require a single-phase program anyway).  A stream
 * that is not an action stream (pop on the bottom builder, unknown register)
 * gives KEX_ERR_ARG, builders nested deeper than 32 - registers
 * KEX_ERR_UNSUPPORTED.
 *
 * Plain pointers and sizes only.  All functions return KEX_OK (0) or a
 * negative KEX_ERR_* code; nothing is printed and nothing calls exit().
 */
#ifndef KEXCUDA_H
#define KEXCUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KEX_OK               0
#define KEX_ERR_BAD_BLOB    -1   /* blob magic/version/offsets invalid            */
#define KEX_ERR_CUDA        -2   /* a CUDA runtime call failed (kex_last_cuda_error) */
#define KEX_ERR_OUT_CAP     -3   /* output buffer too small; *out_len = bytes needed */
#define KEX_ERR_UNSUPPORTED -4   /* program exceeds a device-table limit          */
#define KEX_ERR_ARG         -5   /* bad argument (null, misaligned input, phase)  */
#define KEX_RETRY_EXACT      1   /* kex_shard_emit after kex_set_shard_tail: repeat the shard steps (see there) */

/* Match status, mirroring the exit status of the reference binary
 * (src/KMC/Program/Backends/C.hs:79-81). */
#define KEX_ACCEPT 0
#define KEX_REJECT 1

typedef struct kex_program kex_program;

typedef struct {
  uint32_t nphases;        /* NUM_PHASES of the pipeline (C.hs:41)               */
  uint32_t nstates;        /* SST states of phase 0                              */
  uint32_t nclasses;       /* byte classes of phase 0                            */
  uint32_t nregs;          /* registers incl. the output stream, phase 0         */
  uint32_t nactions;       /* distinct register-update actions, phase 0          */
  uint32_t max_out_per_byte; /* worst-case output bytes per input byte, phase 0  */
  uint32_t chunk_bytes;    /* bytes one device chunk covers                      */
  uint32_t monoid_kernels; /* 1: phase 0 runs on the monoid kernels, 0: generic   */
  uint32_t emit_kernel;    /* emit kernel family the next run of the phase uses: 4 = G-mode (kex_v4.cuh),
                              3 = k3_emit, 2 = CTA-tile monoid kernels, 1 = generic                     */
  uint32_t exact_tiles;    /* G-mode kernel, last run: tiles it had to evaluate exactly                 */
} kex_info_t;

/* Replaces the emit+cc step of compileProgram (C.hs:529-568) and `init()`
 * (C.hs:470-484): parse a kexprog blob and upload its tables to `device`. */
int kex_load(const void *blob, size_t blob_len, int device, kex_program **out);

/* Replaces process teardown; frees tables and scratch. */
void kex_free(kex_program *p);

/* Replaces `-i` / printCompilationInfo (crt/crt.c:384-386, C.hs:49-52). */
int kex_info(const kex_program *p, uint32_t phase, kex_info_t *info);

/* Replaces `run(phase)` for every phase of the pipeline (crt/crt.c:356-364,
 * 414-455) with input and output resident in device memory.  `d_in` and `d_out`
 * must be 16-byte aligned.  On KEX_OK: *status is KEX_ACCEPT or KEX_REJECT, *out_len
 * the bytes written to d_out, *fail_count the reference's `count` (bytes
 * consumed by completed transitions, C.hs:79-81) when rejecting.  A rejecting
 * run leaves exactly the whole 16 KiB flushes the C runtime would have
 * written (crt/crt.c:107-159,217-227).  `stream` is a cudaStream_t or NULL. */
int kex_run_device(kex_program *p, const uint8_t *d_in, size_t n,
                   uint8_t *d_out, size_t out_cap, size_t *out_len,
                   int *status, size_t *fail_count, void *stream);

/* Same, with host buffers: host->device copy of the input, the run, and a
 * device->host copy of the output all happen inside the call
 * (the stdin/stdout role of crt/crt.c:293-312,107-136).  Large inputs of
 * single-phase programs are cut into sub-waves that are copied in, evaluated
 * as consecutive shards of one run and copied out on three streams; pinned
 * host buffers make the copies asynchronous, pageable ones work too. */
int kex_run_host(kex_program *p, const uint8_t *h_in, size_t n,
                 uint8_t *h_out, size_t out_cap, size_t *out_len,
                 int *status, size_t *fail_count);

/* ---- block streaming ------------------------------------------------------
 * Replaces the sliding input window and the flushed output window of the
 * compiled binary (crt/crt.c:61,299-305 readnext; 107-136 buf_flush): the input
 * is fed block by block and never has to be resident as a whole.
 *   kex_stream_begin : start a run
 *   kex_stream_feed  : feed the next input block (host memory); writes to h_out
 *                      the output of the PREVIOUS block (it needs one block of
 *                      lookahead), *out_len = 0 on the first call
 *   kex_stream_end   : output of the last block and of the end-of-input action,
 *                      status and count of the whole run.  On KEX_REJECT the
 *                      caller keeps only whole 16 KiB flushes of the
 *                      concatenated stream, as the reference does.
 * Single-phase programs on the v3 kernels.  A block whose seam summary is not
 * a constant map (e.g. a short last block) keeps its predecessor waiting too;
 * a third block in that situation gives KEX_ERR_UNSUPPORTED -- feed larger
 * blocks or use kex_run_host.  KEX_ERR_OUT_CAP: *out_len = bytes needed; the call changed
 * nothing, repeat it with a larger h_out (same block for kex_stream_feed). */
int kex_stream_begin(kex_program *p);
int kex_stream_feed(kex_program *p, const uint8_t *h_in, size_t n, uint8_t *h_out,
                    size_t out_cap, size_t *out_len);
int kex_stream_end(kex_program *p, uint8_t *h_out, size_t out_cap, size_t *out_len,
                   int *status, size_t *fail_count);

/* ---- sharded evaluation (one shard per GPU; single-phase programs) --------
 * The transducer run is a prefix computation over the SST's transition monoid
 * (src/KMC/SymbolicSST.hs:122-136): forward over the state maps, backward
 * over which registers still reach the output stream.  A shard is evaluated
 * in three steps so that ranks can exchange the two small summaries (one
 * all-gather each) between them:
 *
 *   kex_shard_summarize : state map of the shard for every start state
 *   kex_shard_walk      : given the true start state -> end state, first
 *                         failure, the shard's seam summary (opaque,
 *                         kex_seam_bytes() bytes: how liveness of the
 *                         registers at the shard's end maps to its start)
 *   kex_stitch_live     : all seam summaries + the end-of-input action ->
 *                         seam code at the end of every shard (host only)
 *   kex_shard_emit      : given its seam code -> write the shard's output
 *
 * State maps have nstates+1 uint16 entries (the last is the FAIL sink).     */
int kex_shard_summarize(kex_program *p, const uint8_t *d_in, size_t n,
                        uint16_t *h_state_map, void *stream);
size_t kex_seam_bytes(const kex_program *p);
int kex_shard_walk(kex_program *p, uint32_t start_state, uint32_t *end_state,
                   size_t *fail_pos /* (size_t)-1 if none */, uint8_t *h_seam,
                   void *stream);
int kex_stitch_live(const kex_program *p, const uint8_t *seams, size_t nshards,
                    uint32_t final_code, uint32_t *codes);
int kex_shard_emit(kex_program *p, uint32_t seam_code, size_t n_eff,
                   uint8_t *d_out, size_t out_cap, size_t *out_len, void *stream);

/* Opt-in for the sharded entry points: let kex_shard_walk / kex_shard_emit use the G-mode tail
 * evaluation (exact live sets only for the last tiles of the shard; programs with state-determined
 * live sets, after a first evaluation has learnt them).  kex_shard_walk then reports the shard's seam
 * summary under the assumption that every tile before the tail is consistent with the learnt table;
 * kex_shard_emit verifies it and returns KEX_RETRY_EXACT (1) if it does not hold.  In that case the
 * summary this rank gave to the others was not reliable: EVERY rank must repeat kex_shard_walk ->
 * exchange -> kex_stitch_live -> kex_shard_emit (the failing rank evaluates exactly from then on),
 * so the callers all-reduce the return code of kex_shard_emit.  Off by default: without it the three
 * steps never need repeating. */
int kex_set_shard_tail(kex_program *p, int enabled);

/* Final-state action: is `state` accepting, the seam code of the end-of-input
 * action (which registers it flushes), and its literal tail
 * (SSTCompiler.hs:147-154). */
int kex_final_action(const kex_program *p, uint32_t state, int *accepting,
                     uint32_t *seam_code, const uint8_t **tail, size_t *tail_len);

/* Replaces `-p N` / `--phase N` of the compiled binary (crt/crt.c:380-399,
 * 356-364): the following kex_run_* calls evaluate only phase `phase`
 * (1-based, as the reference numbers them); 0 = the whole pipeline (default). */
int kex_select_phase(kex_program *p, uint32_t phase);

/* Upper bound on the output size for n input bytes (all phases). */
size_t kex_out_bound(const kex_program *p, size_t n);

/* Kernel launches issued by the most recent kex_run_* / kex_shard_* call. */
uint32_t kex_last_launch_count(const kex_program *p);

/* Duration in ms of the dominant (emit) kernel in the most recent run, timed
 * with CUDA events on the launching stream; 0 if timing is disabled. */
int kex_set_timing(kex_program *p, int enabled);
float kex_last_kernel_ms(const kex_program *p, uint32_t which);

const char *kex_strerror(int code);
const char *kex_last_cuda_error(const kex_program *p);

#ifdef __cplusplus
}
#endif
#endif /* KEXCUDA_H */
