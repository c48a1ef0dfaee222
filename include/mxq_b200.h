/*
 * mxq_b200 -- C ABI of the B200-native (sm_100a) MXQ quantization hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types, no C++ exceptions.
 * Every entry point replaces one reference interface (cited as file:line under /root/reference).
 *
 * Conventions
 *   - all tensor pointers are DEVICE pointers on the current CUDA device, row-major, contiguous,
 *     16-byte aligned; the caller owns every buffer (inputs, outputs, workspace) -- the library
 *     never allocates, frees or synchronises;
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *   - return value: 0 = launched; < 0 = argument error (MXQ_E_*); > 0 = cudaError_t of the launch;
 *   - dtype: MXQ_F32 / MXQ_F16 / MXQ_BF16;
 *   - group_bits: optional device uint8[cols/group]; low 7 bits = bit-width, bit 7 (0x80) = the
 *     group belongs to the per-row pool.  NULL = the reference recipe
 *     {low, low, low, 0x80|4} repeated (utils_quant.py:340-385, mxqgpt.py:404-419).
 *   - re-entrant and thread-safe: no mutable globals.
 */
#ifndef MXQ_B200_H_
#define MXQ_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MXQ_VERSION 100

#if defined(__GNUC__)
#define MXQ_API __attribute__((visibility("default")))
#else
#define MXQ_API
#endif

enum { MXQ_F32 = 0, MXQ_F16 = 1, MXQ_BF16 = 2 };

enum {
  MXQ_OK = 0,
  MXQ_E_NULL = -1,      /* required pointer is NULL */
  MXQ_E_SHAPE = -2,     /* shape / divisibility requirement violated */
  MXQ_E_DTYPE = -3,     /* unknown dtype */
  MXQ_E_ALIGN = -4,     /* pointer not 16-byte aligned */
  MXQ_E_UNSUPPORTED = -5,
  MXQ_E_WORKSPACE = -6  /* workspace too small */
};

MXQ_API int mxq_version(void);
MXQ_API const char* mxq_error_string(int code);

/* ---- (a-1) MXAsymQuantizer.forward   LLM-QAT/models/utils_quant.py:315-462 -------------------
 * out[r,c] = round(((x-beta)/(alpha+1e-8))*s)/s*(alpha+1e-8)+beta, every op rounded to `dtype`.
 * codes (optional, uint8[rows*cols]) receives the integer codes round(...) for parity checks.
 * Requires cols % group == 0, group a power of two with 16 bytes <= group*sizeof(dtype) <= 512.
 * With group_bits == NULL also cols % (4*group) == 0. */
MXQ_API int mxq_fakequant_fwd(const void* x, void* out, uint8_t* codes, int64_t rows, int64_t cols,
                      int dtype, int group, int low_bits, const uint8_t* group_bits, void* stream);

/* Row-resident variant for several weights at once and / or an importance mask:
 *  - x / out / rows are HOST arrays of n entries; all tensors have `cols` columns and one dtype.  One
 *    launch covers up to 8 tensors (QAT fake-quantizes the 4 + 2 same-width linears of a decoder layer
 *    back to back; a 67 MB launch alone spends a fifth of its time in ramp and tail).
 *  - pooled_mask (optional, uint8[cols / group]): only the POOL flag (0x80) is read -- WHICH groups share
 *    the row's 4-bit statistic (mxq_allocate_bits); every other group is `low_bits` wide.  NULL = the
 *    positional recipe.  Exactly the arithmetic of mxq_fakequant_fwd (bit-identical outputs).
 *  group 16 or 128; rows of at most 6144 16-byte chunks; 16-bit dtypes with low_bits == 2. */
MXQ_API int mxq_fakequant_fwd_multi(const void* const* x, void* const* out, const int64_t* rows, int n, int64_t cols,
                                    int dtype, int group, int low_bits, const uint8_t* pooled_mask, void* stream);

/* ---- (a-2) MXAsymQuantizer.backward   utils_quant.py:464-475 ---------------------------------
 * grad_in = grad_out; grad_in[x >= hi] = 0; grad_in[x <= lo] = 0.   n = number of elements. */
MXQ_API int mxq_ste_bwd(const void* grad_out, const void* x, void* grad_in, int64_t n, int dtype,
                float lo, float hi, void* stream);

/* ---- (f-1) SymQuantizer / AsymQuantizer.forward   LLM-QAT/models/utils_quant.py:31-95,98-199 --
 * Activation / KV-cache fake quantizers (call sites utils_quant.py:717-721,
 * modeling_llama_quant.py:323-329).  The tensor is `nseg` contiguous segments of `seglen`
 * elements, one statistic per segment (|x|max for mode 0 = symmetric, min/max for mode 1 =
 * asymmetric), every op rounded to `dtype` like the reference's eager ATen chain:
 *   sym : s = reciprocal(m + 1e-6) * (2^(bits-1) - 1);  out = round(x * s) / (s + 1e-6)
 *   asym: a = (max - min) + 1e-8;  out = round(((x - min) / a) * (2^bits - 1)) / (2^bits - 1) * a + min
 * Segment i is live iff (i % period) < valid; dead segments use the zero statistic (the
 * reference's 3-D branch slices dim 1 with column-group indices, :56-64,:144-157, so tokens
 * beyond (C // G) * G keep zero-initialised statistics).  period = 1 disables this.
 * seglen * sizeof(dtype) % 16 == 0.  The backward is mxq_ste_bwd (same clipped STE, :89-95).
 * workspace: mxq_segquant_workspace_bytes(nseg, seglen, dtype) (0 for short segments). */
MXQ_API size_t mxq_segquant_workspace_bytes(int64_t nseg, int64_t seglen, int dtype);
MXQ_API int mxq_segquant_fwd(const void* x, void* out, int64_t nseg, int64_t seglen, int dtype,
                             int mode, int bits, int64_t period, int64_t valid, void* workspace,
                             size_t workspace_bytes, void* stream);

/* ---- (a-5/a-8) calibration statistics --------------------------------------------------------
 * sumsq[c] (+)= sum_t X[t,c]^2 in fp32.  This is the whole of what MXQGPT.add_batch's K x K
 * Hessian is used for (diag(H)==0, mxqgpt.py:369-383,399-403) and the Wanda statistic of
 * WrappedGPT.add_batch (layerwrapper.py:22-35: scaler_row = scaler_row*n/(n+b) + sumsq/(n+b)).
 * out = prev_scale * out + add_scale * sumsq  when accumulate != 0, else out = add_scale*sumsq.
 * `workspace` needs mxq_colsumsq_workspace_bytes(tokens, cols) bytes. */
MXQ_API size_t mxq_colsumsq_workspace_bytes(int64_t tokens, int64_t cols);
MXQ_API int mxq_colsumsq(const void* X, int64_t tokens, int64_t cols, int dtype, float* out,
                 float prev_scale, float add_scale, int accumulate, void* workspace,
                 size_t workspace_bytes, void* stream);

/* Same with an occupancy cap: at most ctas_per_sm (1..8) CTAs of the streaming kernel per SM, so a
 * compute-bound kernel on another stream (the previous layer's quantize+pack) can share the SMs
 * while this one saturates HBM.  ctas_per_sm = 8 is mxq_colsumsq.  Same workspace size. */
MXQ_API int mxq_colsumsq_ex(const void* X, int64_t tokens, int64_t cols, int dtype, float* out,
                    float prev_scale, float add_scale, int accumulate, int ctas_per_sm, void* workspace,
                    size_t workspace_bytes, void* stream);

/* Wanda metric |W| * sqrt(scaler_row)   mxq_quant/lib/prune.py:177.   W: dtype [rows, cols],
 * scaler_row: fp32[cols], out: fp32 [rows, cols]. */
MXQ_API int mxq_wanda_metric(const void* W, const float* scaler_row, float* out, int64_t rows,
                     int64_t cols, int dtype, void* stream);

/* ---- (f-3) importance-driven 2/4-bit allocation ------------------------------------------------
 * importance[g] = sum_{r, c in group g} |W[r,c]| * sqrt(scaler_row[c])   (the Wanda metric of
 * prune.py:177 summed per column group; scaler_row == NULL: plain |W|).  Of every 4 consecutive
 * groups the most important one gets {0x80 | 4} (pooled 4-bit), the others `low_bits`; ties go to
 * the lowest index.  The result is a group_bits mask for mxq_fakequant_fwd / mxq_ptq_quant.
 * W fp16 [rows, cols]; cols % (4*group) == 0; importance (optional) double[cols/group].
 * Deterministic and exact: |W| column sums are accumulated as 64-bit integers in units of 2^-24.
 * workspace: mxq_allocate_bits_workspace_bytes(cols). */
MXQ_API size_t mxq_allocate_bits_workspace_bytes(int64_t cols);
MXQ_API int mxq_allocate_bits(const void* W, const float* scaler_row, int64_t rows, int64_t cols, int group,
                              int low_bits, uint8_t* group_bits, double* importance, void* workspace,
                              size_t workspace_bytes, void* stream);

/* ---- (a-6/a-7) MXQGPT.fasterquant(blocksize=16) + Quantizer ----------------------------------
 * mxq_quant/lib/mxqgpt.py:387-448, mxq_quant/lib/quantizer.py:5-20,61-121,149-155.
 * W (fp16 [rows, cols]) -> Wq (fp16 fake-quantized, may alias W).  colstat (optional fp32[cols]):
 * columns with colstat == 0 are "dead" and zeroed first (mxqgpt.py:401-403).
 * codes (optional uint8 [rows, cols]) receives the integer codes.  rows % 16 == 0 (second-level
 * scale quantisation spans 16 consecutive rows, quantizer.py:115).
 * workspace: mxq_ptq_workspace_bytes(rows, cols). */
MXQ_API size_t mxq_ptq_workspace_bytes(int64_t rows, int64_t cols);
MXQ_API int mxq_ptq_quant(const void* W, void* Wq, uint8_t* codes, const float* colstat, int64_t rows,
                  int64_t cols, int group, int low_bits, const uint8_t* group_bits,
                  void* workspace, size_t workspace_bytes, void* stream);

/* Generic Quantizer(bits, perchannel, asym, qq_scale_bits=4).find_params + quantize_dequantize on
 * an fp32 matrix x[rows, cols] with one scale/zero per row (quantizer.py:61-121,149-155).
 * y, codes, scale, zero are optional outputs (fp32[rows*cols], uint8[rows*cols], fp32[rows] x2). */
MXQ_API int mxq_rowquant(const float* x, float* y, uint8_t* codes, float* scale, float* zero,
                 int64_t rows, int64_t cols, int bits, int qq_scale_bits, void* stream);

/* ---- (a-9) packed mixed 2/4-bit layout -------------------------------------------------------
 * mxq_quant/cuda_kernel/csrc/quantization/gemv_mxq_cuda.cu:39-208 is the only consumer in the
 * reference; shapes for [OC, IC]:
 *   weight int32[OC, IC/16], weight_last int32[OC, IC/64], zeros_and_scales int32[OC, 32*nch],
 *   zeros_2nd int32[OC/4, 32*nch], scales_2nd fp16[OC/4, 3*IC/64], scales_4b fp16[OC],
 *   zeros_4b int32[OC/8];   nch = ceil(IC/4096).  IC % 64 == 0, OC % 8 == 0 (mxq_pack: OC % 16). */
typedef struct {
  int32_t* weight;
  int32_t* weight_last;
  int32_t* zeros_and_scales;
  int32_t* zeros_2nd;
  void* scales_2nd;   /* fp16 */
  void* scales_4b;    /* fp16 */
  int32_t* zeros_4b;
} mxq_packed_t;

/* Quantize fp16 W[OC, IC] into the packed layout (encode policy documented in DESIGN.md; the
 * reference has no producer).  colstat as in mxq_ptq_quant.  Every word of every output tensor
 * is written (no pre-zeroing needed). */
MXQ_API size_t mxq_pack_workspace_bytes(int64_t OC, int64_t IC);
MXQ_API int mxq_pack(const void* W, const float* colstat, int64_t OC, int64_t IC, mxq_packed_t out,
             void* workspace, size_t workspace_bytes, void* stream);

/* Both of the above in one pass over W (read once; writes Wq and the packed tensors): what
 * MXQGPT.fasterquant does per linear plus the packing the reference leaves undone.
 * workspace: mxq_ptq_workspace_bytes(rows, cols). */
MXQ_API int mxq_ptq_quant_pack(const void* W, void* Wq, uint8_t* codes, const float* colstat,
                               int64_t rows, int64_t cols, mxq_packed_t out, void* workspace,
                               size_t workspace_bytes, void* stream);

/* Dequantize the packed layout to fp16 or fp32 [OC, IC] (decode formula of
 * gemv_mxq_cuda.cu:131-136,152-153,178-179,191-192). */
MXQ_API int mxq_unpack(mxq_packed_t in, int64_t OC, int64_t IC, void* out, int out_dtype, void* stream);

/* ---- (a-9) gemv_mxq_forward_cuda   gemv_mxq_cuda.cu:225-273, gemv_mxq_cuda.h:4-12 -------------
 * y[b, oc] = sum_k dequant(W)[oc, k] * x[b, k]; x fp16 [B, IC], y fp16 [B, OC], fp32 accumulate.
 * Any IC % 64 == 0 (the reference is hard-wired to 4096), any B >= 1. */
MXQ_API int mxq_gemv(const void* x, mxq_packed_t w, void* y, int64_t B, int64_t IC, int64_t OC,
             void* stream);

/* Same, with flags.  By default the kernel is launched with the programmatic-dependent-launch
 * attribute: it may become resident while the previous kernel of `stream` is still running.  Until
 * that kernel has completed it only reads the PACKED WEIGHT tensors (TMA bulk copies into shared
 * memory); x is read and y written strictly afterwards.  The overlap is therefore safe unless the
 * previous kernel of the stream writes `w` AND signals early completion itself
 * (griddepcontrol.launch_dependents) -- no kernel of this library does.  MXQ_GEMV_NO_PDL launches
 * fully serialised. */
#define MXQ_GEMV_NO_PDL 1u
MXQ_API int mxq_gemv_ex(const void* x, mxq_packed_t w, void* y, int64_t B, int64_t IC, int64_t OC,
                        unsigned flags, void* stream);

/* Grouped decode GEMV (extension): up to 4 packed linears of the SAME [OC, IC] that share the
 * activation x -- q/k/v, gate/up of a decoder layer -- in one launch: y[i] = x @ dequant(w[i])^T.
 * `w` and `y` are HOST arrays of n entries.  One launch pays the fixed latencies (prologue,
 * activation staging, first weight fill) once per group instead of once per linear. */
MXQ_API int mxq_gemv_grouped(const void* x, const mxq_packed_t* w, void* const* y, int n, int64_t B,
                             int64_t IC, int64_t OC, unsigned flags, void* stream);

/* Decode-GEMV CHAIN (extension): a whole sequence of batch-1 gemv_mxq_forward_cuda calls
 * (gemv_mxq_cuda.cu:225-273 launches one kernel per linear) as ONE persistent launch.  The packed
 * weights of all jobs stream through shared memory back to back -- a linear boundary costs no launch,
 * no fill and no drain -- which is what lets a 6 MB GEMV approach the HBM rate.
 *   job j:  y_j[OC_j] = dequant(w_j) @ x_j[IC_j]      (fp16 in / out, fp32 accumulate, batch 1)
 *   dep = -1: x_j is complete when the launch starts.  dep = i < j: x_j is (or depends on) y_i -- the
 *   kernel orders the read of x_j after every store of y_i.  Jobs are otherwise UNORDERED: a job must
 *   not write a buffer that an unordered job reads or writes.
 * Requirements: IC % 256 == 0, IC <= 32768, OC % 32 == 0, n <= MXQ_GEMV_CHAIN_MAX_JOBS (split longer
 * chains; stream order covers dependencies between launches); MXQ_E_UNSUPPORTED otherwise -- callers fall
 * back to mxq_gemv per linear.
 *   mxq_gemv_chain_plan   validates the HOST job array and writes a plan of mxq_gemv_chain_plan_bytes()
 *                         bytes to HOST memory (64-byte aligned): the job table, which is passed to the
 *                         kernel by value, and four TMA tensor maps per job.  The caller keeps the host
 *                         plan and a DEVICE copy of it (64-byte aligned; the kernel reads the tensor maps
 *                         from there).  Needs a current CUDA context (cuTensorMapEncodeTiled).
 *   mxq_gemv_chain_run    launches the chain.  sync_ws: device int32[MXQ_GEMV_CHAIN_SYNC_WORDS], zeroed ONCE
 *                         by the caller; the kernel re-arms it at exit (graph replays need no memset).
 *                         Chains with dependencies are launched cooperatively (all CTAs co-resident).
 *                         flags: MXQ_GEMV_CHAIN_PDL -- programmatic dependent launch for chains without
 *                         dependencies: the kernel builds its tile lists and prefetches packed weights while
 *                         the PREVIOUS kernel of the stream is still draining (its CTAs leave the SMs one by
 *                         one), and waits for that kernel before it reads an activation vector or writes an
 *                         output.  Contract: the previous kernel does not write this chain's packed weights
 *                         or its plan.  A preceding chain launch lets its successor in at once; any other
 *                         kernel is waited for as usual. */
#define MXQ_GEMV_CHAIN_MAX_JOBS 64
#define MXQ_GEMV_CHAIN_PDL 1u
#define MXQ_GEMV_CHAIN_SYNC_WORDS (MXQ_GEMV_CHAIN_MAX_JOBS + 1)
typedef struct {
  const void* x;      /* fp16 [IC] */
  void* y;            /* fp16 [OC] */
  mxq_packed_t w;
  int64_t IC, OC;
  int32_t dep;        /* -1, or the index of an earlier job whose y this job's x depends on */
  int32_t reserved;
} mxq_gemv_job_t;
MXQ_API size_t mxq_gemv_chain_plan_bytes(void);
MXQ_API int mxq_gemv_chain_plan(const mxq_gemv_job_t* jobs, int n, void* plan_host);
MXQ_API int mxq_gemv_chain_run(const void* plan_host, const void* plan_dev, int32_t* sync_ws, unsigned flags,
                               void* stream);

/* ---- (f-3) importance-driven allocation folded into the packed path ----------------------------------
 * The packed layout is positional (the last 16 of every 64 columns are the 4-bit ones,
 * utils_quant.py:349-353, mxqgpt.py:404-419).  A data-driven choice of the 4-bit group (mxq_allocate_bits,
 * Wanda metric prune.py:177; act-order permutation weight_permutation.py:27-93) is expressed as a
 * permutation of 16-column groups: packed group g holds original group group_perm[g].  The weights are
 * gathered once before packing (mxq_gather_groups), the prefill GEMM gathers its activations the same way,
 * and the decode GEMV applies the permutation while it stages x in shared memory (no extra pass). */
MXQ_API int mxq_gather_groups(const void* in, const int32_t* group_perm, void* out, int64_t rows, int64_t cols,
                              void* stream);   /* fp16 [rows, cols], cols % 16 == 0, out != in */
MXQ_API int mxq_gemv_grouped_perm(const void* x, const mxq_packed_t* w, void* const* y, int n, int64_t B,
                                  int64_t IC, int64_t OC, const int32_t* group_perm /* int32[IC/16] or NULL */,
                                  unsigned flags, void* stream);

/* ---- (a-10) gemv_forward_cuda (AWQ uniform 4-bit)   gemv_cuda.cu:346-399, gemv_cuda.h:4-9 ------
 * kernel int32[OC, IC/8] (nibble j of word i = column 8i+j), zeros int32[OC, zw] (nibble g%8 of
 * word g/8, g = col/G), scales fp16[OC, zw*8]; zw = ceil(IC/G/8) rounded up to 1/2/4 words for
 * G = 128/64/32 (gemv_cuda.cu:200,129,56). */
MXQ_API int mxq_awq_gemv(const void* x, const int32_t* kernel, const void* scales, const int32_t* zeros,
                 void* y, int64_t B, int64_t IC, int64_t OC, int group_size, void* stream);

/* ---- AWQ uniform 4-bit prefill GEMM   gemm_cuda_gen.cu:424-478 (declared in gemm_cuda.h, never built by
 * the reference's setup.py:37-41) ------------------------------------------------------------------
 * y[m, n] = sum_k x[m, k] * W[k][n],  W[k][n] = fp16(scales[k/G][n] * (q - z)).  x fp16 [M, IC];
 * kernel int32 [IC, OC/8], zeros int32 [IC/G, OC/8]: output channel 8j + c = nibble
 * {0,4,1,5,2,6,3,7}[c] of word j (dequantize.cuh:15-77); scales fp16 [IC/G, OC]; y fp16 [M, OC].
 * G % 32 == 0, OC % 64 == 0, OC % G == 0 (the reference's own checks, :447-454), IC % 64 == 0.
 * fp32 accumulation over all of K (the reference's `split_k_iters` fp16 partial sums do not exist).
 * workspace: mxq_awq_gemm_workspace_bytes(IC, OC) (the dequantized fp16 operand). */
MXQ_API size_t mxq_awq_gemm_workspace_bytes(int64_t IC, int64_t OC);
MXQ_API int mxq_awq_gemm(const void* x, const int32_t* kernel, const void* scales, const int32_t* zeros, void* y,
                         int64_t M, int64_t IC, int64_t OC, int group_size, void* workspace,
                         size_t workspace_bytes, void* stream);

/* ---- prefill: packed dequant-GEMM on tcgen05/TMEM (no reference kernel exists for the mixed
 * layout; gemm_cuda_gen.cu:424-478 is the un-built AWQ 4-bit analogue) --------------------------
 * y[m, oc] = sum_k x[m, k] * dequant(W)[oc, k]; x fp16 [M, IC], y fp16 [M, OC].
 * workspace: mxq_gemm_workspace_bytes(M, IC, OC), uninitialised.  When the tiles (512 tokens x 256
 * rows) do not fill a whole number of waves of SM pairs, the tiles of the last wave are cut along K
 * and their fp32 partials meet in the workspace (a second small kernel adds them in a fixed order:
 * results are reproducible).  A NULL or too small workspace is legal: whole tiles only.  Calls that
 * may run concurrently (different streams) need separate workspaces. */
MXQ_API size_t mxq_gemm_workspace_bytes(int64_t M, int64_t IC, int64_t OC);
MXQ_API int mxq_gemm(const void* x, mxq_packed_t w, void* y, int64_t M, int64_t IC, int64_t OC,
             void* workspace, size_t workspace_bytes, void* stream);
/* The tile schedule mxq_gemm would use on a device with `sms` SMs (pure host arithmetic, no device
 * needed): out5 = {M tiles, N tiles, tiles, whole tiles, K slices per remaining tile}. */
MXQ_API int mxq_gemm_plan(int64_t M, int64_t IC, int64_t OC, int sms, int32_t* out5, size_t* workspace_bytes);

/* Column-sharded GEMM fused with its all-gather: this rank owns output columns
 * [col0, col0 + OC) of a [M, ldy] result; the epilogue stores every tile into each of the `npeers`
 * output buffers in y_peers (host array of device pointers: the local buffer and the peers'
 * buffers mapped into this process over NVLink, e.g. torch symmetric memory).  The caller
 * synchronises the ranks afterwards (the library never does).  workspace as for mxq_gemm (the
 * second pass of K-split tiles then carries the exchange stores); NULL is legal. */
MXQ_API int mxq_gemm_scatter(const void* x, mxq_packed_t w, void* const* y_peers, int npeers,
                             int64_t M, int64_t IC, int64_t OC, int64_t ldy, int64_t col0,
                             void* workspace, size_t workspace_bytes, void* stream);

/* Same exchange through an NVSwitch multicast mapping: `y_multicast` is the multicast address of a
 * symmetric [M, ldy] fp16 buffer (e.g. torch symmetric memory `multicast_ptr`); the epilogue issues
 * ONE multimem.st per 16 bytes and the switch replicates it into every rank's buffer, so a rank's
 * NVLink egress is its tile once instead of once per peer.  The caller synchronises the ranks
 * afterwards.  MXQ_E_UNSUPPORTED is never returned here: whether the address is a multicast
 * mapping is the caller's contract. */
MXQ_API int mxq_gemm_multicast(const void* x, mxq_packed_t w, void* y_multicast, int64_t M, int64_t IC,
                               int64_t OC, int64_t ldy, int64_t col0, void* workspace,
                               size_t workspace_bytes, void* stream);

/* Phased exchange for shards that are a single partly filled wave (70B shapes at 8 ranks): per group of
 * weight rows, (1) the GEMM with EVERY tile cut into `split` K slices -- fp32 partials into the
 * workspace, no output written -- and (2) the second pass alone: slices added in slice order, fp16 rows
 * stored into the peers' buffers (y_peers, npeers) or once into the multicast mapping (y_multicast, then
 * y_peers may be NULL).  Run (2) of group g on another stream than (1) of group g+1 and the exchange
 * overlaps the tensor work (mxq_b200/dist.py, modes "p2p2" / "mc2").  IC % 256 == 0,
 * 1 <= split <= IC / 256.  workspace: mxq_gemm_partials_workspace_bytes(M, OC, split). */
MXQ_API size_t mxq_gemm_partials_workspace_bytes(int64_t M, int64_t OC, int split);
MXQ_API int mxq_gemm_partials(const void* x, mxq_packed_t w, int64_t M, int64_t IC, int64_t OC, int split,
                              void* workspace, size_t workspace_bytes, void* stream);
MXQ_API int mxq_gemm_reduce_store(const void* workspace, size_t workspace_bytes, void* const* y_peers, int npeers,
                                  void* y_multicast, int64_t M, int64_t OC, int split, int64_t ldy, int64_t col0,
                                  void* stream);

/* Diagnostic: the same tcgen05/TMA pipeline with a dense fp16 B operand W[OC, IC] loaded by TMA
 * instead of dequantized in registers (y = x @ W^T).  Separates UMMA-descriptor errors from
 * dequant/swizzle errors in tests; not part of the reference surface. */
MXQ_API int mxq_gemm_dense(const void* x, const void* W, void* y, int64_t M, int64_t IC, int64_t OC,
                           void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MXQ_B200_H_ */
