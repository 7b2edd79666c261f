/* opp_b200.h — C ABI of libopp_b200.so: the B200 (sm_100a) kernels behind the OnePose++ 2D-3D
 * coarse-to-fine matcher, `OnePosePlus_model.forward` (reference:
 * src/models/OnePosePlus/OnePosePlusModel.py:96-201).
 *
 * The reference has no FFI on this path: every stage is a PyTorch library call.  The entry
 * points below are what a binding for this path would bind — one per reference stage — and each
 * one cites the reference code it replaces.  `onepose_plus_plus_b200/_lib.py` is the ctypes
 * binding; INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in _host; fp16 buffers are `void*`
 *   - no allocation and no synchronisation inside; work is enqueued on `stream`
 *   - return 0 on success, non-zero on error; opp_last_error() describes the last failure of the
 *     calling thread
 *   - feature maps are NHWC fp16 with the channel count padded to a multiple of 16
 *     (196 -> 208); token tensors are [batch][tokens][channels]
 *   - `split`: every fp16 tensor that feeds a tensor-core GEMM is stored as two planes along its
 *     channel axis, row = [hi(C) | lo(C)] with hi = fp16(x), lo = fp16(x - hi); GEMMs then issue
 *     hi*hi + hi*lo + lo*hi into one fp32 accumulator (fp32-grade result, parity mode).  With
 *     split = 0 rows are plain fp16 [C] (2^-11 operands; does not meet the 1e-3 parity bar).
 *     All `ld`/channel counts below are per plane; split = 1 doubles the row length.
 */
#ifndef OPP_B200_H_
#define OPP_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef void* opp_stream_t; /* cudaStream_t */

int opp_version(void);
const char* opp_last_error(void);
int opp_num_sms(void);


/* ------------------------------------------------------------------------------------------
 * Backbone — ResNetFPN_8_2.forward (backbone/resnet.py:141-164), BatchNorm folded on the host;
 * every convolution runs on the tcgen05 engine, the two bilinear x2 upsample-adds of the FPN are
 * epilogues of the lateral 1x1 convolutions
 * ---------------------------------------------------------------------------------------- */

/* conv1 on the tensor-core engine: im2col of the 7x7 stride-2 pad-3 windows (resnet.py:101-103).
 * a_out fp16 [B*H/2*W/2][planes*64]: row = (49 taps, 1.0, 14 zeros) of one output pixel, so that
 * opp_linear_act_f16(a_out, k0 = 64, w = [c_out][planes*64] holding (49 folded-BN taps, folded
 * bias, zeros), act = ReLU) yields the NHWC map [B][H/2][W/2][planes*c_out] of
 * relu(bn1(conv1(x))) (resnet.py:143).  image: fp32 [B][1][H][W] in [0,1], or (image_u8 != 0)
 * uint8 with x = u8 / 255 folded in (the host-side division of data_io.py:107). */
int opp_conv1_im2col(const void* image, int image_u8, void* a_out, int batch, int h, int w, int split,
                     opp_stream_t stream);

/* 3x3 (pad 1) or 1x1 (pad 0) convolution, stride 1 or 2, as a tcgen05 implicit GEMM
 * (resnet.py:10-17 conv1x1/conv3x3; BasicBlock resnet.py:36-45; FPN heads resnet.py:109-124).
 *   in    NHWC fp16 [B][in_h][in_w][c_in_pad]
 *   w     fp16 [c_out_pad][planes][ksize*ksize][c_in_pad]   (BN-folded, zero in the padding)
 *   bias  fp32 [c_out_pad]
 *   resid NHWC fp16 [B][out_h][out_w][c_out_pad] added before the activation, or NULL
 *   act   0 none, 1 ReLU, 2 LeakyReLU(slope)
 *   out   NHWC fp16 [B][out_h][out_w][c_out_pad], or NULL when only tokens are wanted
 *   tok/pe: when tok != NULL also writes  out + pe  as coarse tokens
 *          [B][out_h*out_w][planes*c_out_pad]; pe is fp32 [out_h*out_w][c_out_pad]
 *          (PositionEncodingSine.forward position_encoding.py:37-42 + the 'n c h w -> n (h w) c'
 *          rearrange OnePosePlusModel.py:137-142)
 *   up:    when up != NULL the epilogue adds  bilinear_x2(up), align_corners=True  (FPN top-down
 *          merge, resnet.py:149-157: F.interpolate(..., scale_factor=2, mode="bilinear",
 *          align_corners=True) + lateral conv); up is NHWC fp16 [B][out_h/2][out_w/2][c_out_pad] */
int opp_conv2d_nhwc(const void* in, const void* w, const float* bias, const void* resid,
                    void* out, int batch, int in_h, int in_w, int c_in_pad, int c_out_pad,
                    int ksize, int stride, int act, float slope, void* tok, const float* pe,
                    const void* up, int split, opp_stream_t stream);

/* The same 3x3 / stride 1 / pad 1 convolution evaluated only on a win x win window around each
 * coarse match (win = 7 or 5).  The fine branch of the FPN (resnet.py:155-157 layer1_outconv2) is
 * read only inside the 5x5 window of each match (fine_preprocess.py:40-47: unfold, then keep the
 * matched cells), so for few matches the two half-resolution convolutions run on those windows
 * alone: conv A on 7x7 (what conv B's 5x5 outputs need), conv B on 5x5; values are those of the
 * dense convolution at the same positions.
 *   out   fp16 [matches][win][8][planes*c_out_pad] (compact windows; column 7 of a row is padding;
 *         positions outside the image are written as zeros = the padding the next conv must see)
 *   j_ids != NULL (with b_ids): `in` is the dense NHWC map [batch][in_h][in_w][planes*c_in_pad];
 *         window m starts at (x, y) = (stride * cx + org, stride * cy + org), (cy, cx) = divmod(j_ids[m], wc)
 *   j_ids == NULL: `in` is the compact output [matches][win + 2][8][planes*c_in_pad] of a previous
 *         call; output (ly, lx) reads input rows ly..ly+2, columns lx..lx+2
 *   count: NULL = `matches` is exact; else the capacity, the real count is read on the device */
int opp_conv_win_pitch(int win);   /* row pitch P of the output windows: out is [matches][win][P][planes*c_out_pad] */
int opp_conv_win(const void* in, const void* w, const float* bias, void* out, const long long* b_ids,
                 const long long* j_ids, int matches, const int* count, int batch, int in_h, int in_w,
                 int c_in_pad, int c_out_pad, int win, int wc, int stride, int org, int act,
                 float slope, int split, opp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * 3D keypoint encoding — normalize_3d_keypoints (utils/normalize.py:16-26) +
 * KeypointEncoding_linear.forward (utils/position_encoding.py:54-60)
 * ---------------------------------------------------------------------------------------- */

/* stats fp32 [B][4] = (mean x, mean y, mean z, 0.6 * max extent of batch element 0) */
int opp_kpt_stats(const float* kpts, float* stats, int batch, int n, opp_stream_t stream);

/* tokens = desc^T + MLP(normalised kpts); MLP = Linear,IN,ReLU x3 + Linear with per-point
 * instance norm over channels (eps 1e-5).  wN_t are the transposed weights [in][out].
 * kpts fp32 [B][N][3]; desc fp32 [B][256][N]; tok fp16 [B][N][planes*256] */
int opp_kpt_encode(const float* kpts, const float* stats, const float* desc, const float* w1_t,
                   const float* b1, const float* w2_t, const float* b2, const float* w3_t,
                   const float* b3, const float* w4_t, const float* b4, void* tok, int batch, int n,
                   int split, opp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Transformer — LoFTREncoderLayer.forward (loftr_module/transformer.py:65-94) and
 * LinearAttention.forward (loftr_module/linear_attention.py:29-61), tcgen05 GEMMs
 * ---------------------------------------------------------------------------------------- */

/* out[rows][n] = act(concat_K(a0[rows][k0], a1[rows][k1]) @ w[n][k0+k1]^T), fp16 in/out.
 * act 0 none, 1 ReLU, 2 elu(x)+1, applied to output columns < act_cols.
 * Used for [k_proj;v_proj] (transformer.py:78-79 + linear_attention.py:46), mlp.0 + ReLU
 * (transformer.py:41-45,91) and the fine-level q/k/v projections. a1 may be NULL (k1 = 0). */
int opp_linear_act_f16(const void* a0, int k0, const void* a1, int k1, const void* w, void* out,
                       long long rows, int n, int act, int act_cols, int split,
                       opp_stream_t stream);

/* Batched form: a0 / a1 / out are [batches][rows][..]; with a0_shared != 0 the first operand is
 * [1][rows][k0] — one object's tokens shared by every image of the batch (the 3D side of the
 * first cross layer, whose x is image-independent: transformer.py:148-159) — and is read once.
 * row_mask (uint8 [batches*rows], or NULL): rows whose mask is 0 are written as zeros — padded
 * source positions of query_image_mask (linear_attention.py:51-53: K and V are multiplied by kv_mask). */
int opp_linear_act_f16_b(const void* a0, int k0, int a0_shared, const void* a1, int k1, const void* w,
                         void* out, int batches, long long rows, int n, int act, int act_cols,
                         int split, const unsigned char* row_mask, opp_stream_t stream);

/* Same GEMM with split (hi|lo) operands but a single-plane fp16 output [rows][n]: for the K'/V rows
 * of the linear-attention state, whose consumer sums over thousands of rows (built for the next
 * GPU session; selected by $OPP_B200_KV1). */
int opp_linear_act_f16_out1(const void* a0, int k0, const void* a1, int k1, const void* w, void* out,
                            long long rows, int n, int act, int act_cols, const unsigned char* row_mask,
                            opp_stream_t stream);

/* opp_linear_act_f16 / opp_linear_ln with a device-side row count: rows = *count * rows_per_count
 * (clamped to cap_rows, the size the operands were allocated for). */
int opp_linear_act_f16_dyn(const void* a0, int k0, const void* a1, int k1, const void* w, void* out,
                           long long cap_rows, const int* count, int rows_per_count, int n, int act,
                           int act_cols, int split, opp_stream_t stream);
int opp_linear_ln_dyn(const void* a0, int k0, const void* a1, int k1, const void* w, const float* gamma,
                      const float* beta, float eps, const void* resid, void* out16, float* out32,
                      long long cap_rows, const int* count, int rows_per_count, int n, int split,
                      opp_stream_t stream);

/* q_proj + feature map + normaliser (transformer.py:77, linear_attention.py:45,58):
 * out = Q * v_len / (Q . ksum_head + eps), Q = elu(x @ wq^T) + 1, heads of 32 channels.
 * x fp16 [B][rows][256] (or [1][rows][256] with x_shared != 0); ksum fp32 [B][256];
 * out fp16 [B][rows][256]; row_mask uint8 [B*rows] or NULL: Q = 0 on padded query positions
 * (linear_attention.py:49-50) */
int opp_linear_q_f16(const void* x, const void* wq, const float* ksum, void* out, int batches,
                     int rows, int d_model, float v_len, float eps, int split, int x_shared,
                     const unsigned char* row_mask, opp_stream_t stream);

/* y = LayerNorm(concat_K(a0,a1) @ w^T) [+ resid]  (transformer.py:85-94).
 * w fp16 [n][planes*k] or, when w_batched, [B][n][planes*k] (the per-image
 * blockdiag(KV) @ merge^T  matrix).  resid / out16 fp16 [B*rows][planes*n]; out32 fp32
 * [B*rows][n]; either output may be NULL.  resid_shared != 0: resid is [1][rows][planes*n]. */
int opp_linear_ln(const void* a0, int k0, const void* a1, int k1, const void* w, int w_batched,
                  const float* gamma, const float* beta, float eps, const void* resid,
                  int resid_shared, void* out16, float* out32, int batches, long long rows, int n,
                  int split, opp_stream_t stream);

/* FullAttention.forward (linear_attention.py:64-95): out = softmax(Q K^T / sqrt(head_dim)) V per
 * head; `attention: "full"` in the transformer config (no shipped configuration selects it).
 * q fp16 [B][l][planes*heads*head_dim] = q_proj(x); kv fp16 [B][s][planes*2*heads*head_dim] =
 * (k_proj(source) | v_proj(source)); out like q.  head_dim 32 (coarse) or 16 (fine). */
int opp_full_attention(const void* q, const void* kv, void* out, int batch, int l, int s, int heads,
                       int head_dim, int split, opp_stream_t stream);

/* Source side state of linear attention (linear_attention.py:55-57):
 * kv16 fp16 [B][S][planes*2d] holds K' = elu(k)+1 in columns [0,d) and V in [d,2d) of each plane.
 * part fp32 [B][chunks][H][33][32], chunks = opp_kv_chunks_b(S, B): per-chunk sum_s K'^T V (rows
 * 0..31) and sum_s K' (row 32) for each of the H = d/32 heads.  opp_kv_partial picks the chunk length
 * from (S, B) — 256 tokens, 128 when that would leave fewer than 64 CTAs — and opp_kv_finalize must
 * be given the matching chunk count: opp_kv_chunks_b(S, B).  opp_kv_chunks(S) = the 256-token count
 * (what large batches use), kept for callers that size buffers once. */
int opp_kv_chunks(int s);
int opp_kv_chunks_b(int s, int batch);
int opp_kv_partial(const void* kv16, float* part, int batch, int s, int d, int split,
                   opp_stream_t stream);

/* Reduces the chunk partials, scales KV by 1/v_len and folds the merge projection
 * (transformer.py:85): mt[b][c][h*32+dd] = sum_v merge_w[c][h*32+v] * KV[b][h][dd][v] / v_len.
 * merge_w fp32 [d][d]; mt fp16 [B][d][planes*d]; ksum fp32 [B][d]. */
int opp_kv_finalize(const float* part, const float* merge_w, void* mt, float* ksum, int batch,
                    int chunks, int d, float v_len, int split, opp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Coarse matching — CoarseMatching.forward / get_coarse_match
 * (utils/coarse_matching.py:76-123,125-242)
 * ---------------------------------------------------------------------------------------- */

int opp_sim_tiles(int cols); /* column tiles used by the two calls below */

/* per-row partial (max, sum exp) of sim = scale * a @ b^T over each column tile.
 * a fp16 [B][rows][planes*k], b fp16 [B][cols][planes*k];
 * part_m/part_s fp32 [B*rows][opp_sim_tiles(cols)] */
int opp_sim_lse(const void* a, const void* b, float* part_m, float* part_s, int batches, int rows,
                int cols, int k, float scale, int split, opp_stream_t stream);

/* lse[r] = logsumexp over tiles */
int opp_lse_finalize(const float* part_m, const float* part_s, float* lse, long long rows,
                     int tiles, opp_stream_t stream);

/* conf = exp((2 sim - lse_pt) - lse_px) (coarse_matching.py:115); per-row per-tile (max, first
 * argmax); conf (or NULL) fp32 [B*rows][cols] is data["conf_matrix"] when rows are 3D points. */
int opp_sim_conf(const void* a, const void* b, const float* lse_own, const float* lse_other,
                 int own_is_pt, float* conf, float* part_val, int* part_idx, int batches,
                 int rows, int cols, int k, float scale, int split, opp_stream_t stream);

/* opp_sim_lse for rows = 3D points that also produces the COLUMN statistics (saves the second lse
 * pass): col_m/col_s fp32 [B][ceil(rows/32)][cols] = per 32-row group (max, sum exp(x - max)) of
 * every column; opp_lse_col_finalize merges the groups into lse[b][s] = logsumexp_l sim[b, l, s].
 * col_mask (uint8 [B][cols] or NULL) = query_image_mask at coarse resolution: masked columns get
 * sim - 1e9 (coarse_matching.py:108-114), i.e. they drop out of every row's softmax, and
 * opp_lse_col_finalize writes lse = +inf for them so that conf is exactly 0 there. */
int opp_sim_lse_cols(const void* a, const void* b, float* part_m, float* part_s, float* col_m,
                     float* col_s, int batches, int rows, int cols, int k, float scale, int split,
                     const unsigned char* col_mask, opp_stream_t stream);
int opp_lse_col_finalize(const float* col_m, const float* col_s, float* lse, int batches, int groups,
                         int cols, const unsigned char* col_mask, opp_stream_t stream);

/* opp_sim_conf for rows = 3D points with the column maxima folded in (saves the second conf pass):
 * colmax uint32 [B][cols] receives the float bits of max_l conf[b, l, s] (zeroed inside, then
 * atomicMax per 32-row group; conf >= 0 so the bits order like the values). */
int opp_sim_conf_colmax(const void* a, const void* b, const float* lse_own, const float* lse_other,
                        float* conf, float* part_val, int* part_idx, unsigned* colmax, int batches,
                        int rows, int cols, int k, float scale, int split, opp_stream_t stream);

/* best[r] = max over tiles (ties -> lowest index) */
int opp_best_finalize(const float* part_val, const int* part_idx, float* best_val, int* best_idx,
                      long long rows, int tiles, opp_stream_t stream);

/* Threshold + top/left border + mutual nearest neighbour + ordered compaction
 * (coarse_matching.py:142-172, 223-239).  Capacity of every output is batch*min(l, s).
 *   pt_val/pt_idx [B][l]: row maxima of conf;  px_idx [B][s]: column argmax of conf
 *   kpts fp32 [B][l][3]; img_scale fp32 [B][2] = (h_scale, w_scale) or NULL
 *   scratch int32 [ceil(B*l/1024) + 2]
 *   bank_shared != 0: one object for the whole batch, kpts is [1][l][3] (no per-image copies)
 * Outputs (ascending (b, i) order): b_ids/i_ids/j_ids int64, mconf fp32, mkpts3d fp32 [.][3],
 * mkpts_c fp32 [.][2]; count_out int32 [1] = number of matches. */
int opp_match_select(const float* pt_val, const int* pt_idx, const int* px_idx, const float* kpts,
                     const float* img_scale, int batch, int l, int hc, int wc, float thr,
                     int border, float cell, int* scratch, long long* b_ids, long long* i_ids,
                     long long* j_ids, float* mconf, float* mkpts3d, float* mkpts_c,
                     int* count_out, int bank_shared, opp_stream_t stream);

/* opp_match_select with the mutual-nearest test on values (coarse_matching.py:157-165 compares
 * conf == conf.max(dim) the same way): row l keeps its argmax cell j iff pt_val[b][l] has the same
 * float bits as colmax[b][j] (from opp_sim_conf_colmax). */
int opp_match_select_colmax(const float* pt_val, const int* pt_idx, const unsigned* colmax,
                            const float* kpts, const float* img_scale, int batch, int l, int hc,
                            int wc, float thr, int border, float cell, int* scratch,
                            long long* b_ids, long long* i_ids, long long* j_ids, float* mconf,
                            float* mkpts3d, float* mkpts_c, int* count_out, int bank_shared,
                            opp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fine level — FinePreprocess (loftr_module/fine_preprocess.py:32-55), loftr_fine,
 * FineMatching (utils/fine_matching.py:28-110)
 * ---------------------------------------------------------------------------------------- */

/* Every fine-level entry point takes `count_dev`: NULL = `m` is the exact number of matches (known
 * on the host); otherwise `m` is the CAPACITY the buffers were sized for and the kernels read the
 * real match count from *count_dev (int32, device; written by opp_match_select*), so the whole
 * forward can be enqueued / captured in a CUDA graph without the host learning M first (the
 * reference synchronises in torch.where: coarse_matching.py:170). */

/* For match m: row 26m = descriptors3d_db[b, :, i]; rows 26m+1+ww = the 5x5 window (ww = ky*5+kx)
 * of the fine map centred on fine pixel (stride*jy, stride*jx), zero outside the map.
 * fine NHWC fp16 [B][hf][wf][planes*128]; desc3d fp32 [B][128][n];
 * x32 fp32 [26 M][128] (may be NULL) / x16 fp16 [26 M][planes*128];
 * bank_shared != 0: desc3d is [1][128][n], shared by every batch element;
 * windows != 0: `fine` is not the dense map but the compact per-match windows written by
 * opp_conv_win (win = 5): fp16 [M][5][8][planes*128] */
int opp_fine_gather(const void* fine, const float* desc3d, const long long* b_ids,
                    const long long* i_ids, const long long* j_ids, float* x32, void* x16, int m,
                    int hf, int wf, int wc, int stride, int n, int split, int bank_shared,
                    int windows, const int* count_dev, opp_stream_t stream);

/* Linear attention for the 1 + 25 tokens of each match (linear_attention.py:29-61 with
 * L,S in {1,25}).  qkv fp16 [26 M][planes*384] = (elu(q)+1 | elu(k)+1 | v), 8 heads of 16.
 * cross = 0: self layer (each sequence attends to itself); 1: cross layer, both directions from
 * the pre-update tensors (transformer.py:154-159).  msg fp16 [26 M][planes*128] */
int opp_fine_attention(const void* qkv, void* msg, int m, int cross, float eps, int split,
                       const int* count_dev, opp_stream_t stream);

/* Correlation softmax + expectation + std (fine_matching.py:78-94) and sub-pixel coordinates
 * (fine_matching.py:96-110).  x32 fp32 [26 M][128]; img_scale fp32 [B][2] or NULL;
 * expec_f fp32 [M][3]; mkpts_f fp32 [M][2] */
int opp_fine_match(const float* x32, const float* mkpts_c, const long long* b_ids,
                   const float* img_scale, float* expec_f, float* mkpts_f, int m, float fine_scale,
                   const int* count_dev, opp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * LoFTR 2D-2D matcher — LoFTR_for_OnePose_Plus.forward
 * (src/KeypointFreeSfM/loftr_for_sfm/loftr.py:35-127; modules of submodules/LoFTR/src/loftr).
 * Backbone, transformer layers and dual-softmax passes are the entry points above; these four
 * are what differs from the 2D-3D matcher.
 * ---------------------------------------------------------------------------------------- */

/* LoFTR get_coarse_match (utils/coarse_matching.py:133-259, inference): threshold, `border` cells
 * removed on ALL sides of BOTH grids (:9-28), mutual nearest neighbour by value (rowmax == colmax),
 * ordered compaction.  pt_val/pt_idx [B][h0*w0] row maxima / argmax over image 1's cells, colmax
 * [B][h1*w1] (from opp_sim_conf_colmax).  scale0/scale1 fp32 [B][2] or NULL multiply (x, y) as
 * given (:248-253).  Capacity of the outputs: B*h0*w0.  scratch int32 [ceil(B*h0*w0/1024) + 2]. */
int opp_match_select_2d(const float* pt_val, const int* pt_idx, const unsigned* colmax, const float* scale0,
                        const float* scale1, int batch, int h0, int w0, int h1, int w1, float thr,
                        int border, float cell, int* scratch, long long* b_ids, long long* i_ids,
                        long long* j_ids, float* mconf, float* mkpts0_c, float* mkpts1_c, int* count_out,
                        opp_stream_t stream);

/* LoFTR FinePreprocess (loftr_module/fine_preprocess.py:30-59, fine_concat_coarse_feat False):
 * window x window patches (zero outside the map) of both fine maps, sequence-major rows
 * (seq * m + match) * window^2 + ww; seq 0 = image 0 centred on cell i_ids, seq 1 = image 1 on j_ids.
 * fine0/fine1 NHWC fp16 [B][hf][wf][planes*128]; x16 fp16 [2*m*window^2][planes*128]. */
int opp_fine_gather_2d(const void* fine0, const void* fine1, const long long* b_ids, const long long* i_ids,
                       const long long* j_ids, void* x16, int m, int hf0, int wf0, int wc0, int hf1, int wf1,
                       int wc1, int stride, int window, int split, opp_stream_t stream);

/* LinearAttention.forward (linear_attention.py:29-61) between small token groups, 8 heads x 16:
 * group g: q fp16 [g][l][planes*128] = elu(q_proj x)+1, kv fp16 [g][s][planes*256] =
 * (elu(k_proj src)+1 | v_proj src); out like q. */
int opp_seq_attention(const void* q, const void* kv, void* out, int groups, int l, int s, float eps, int split,
                      opp_stream_t stream);

/* LoFTR FineMatching.forward (utils/fine_matching.py:17-74): x32 fp32 [2][m][window^2][128]
 * (sequence-major), centre token of seq 0 against seq 1; expec_f [m][3], mkpts1_f [m][2] =
 * mkpts1_c + coords * (window // 2) * fine_scale * scale1[b]. */
int opp_fine_match_2d(const float* x32, const float* mkpts1_c, const long long* b_ids, const float* scale1,
                      float* expec_f, float* mkpts1_f, int m, int window, float fine_scale,
                      opp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Pose from the matches — ransac_PnP (src/utils/metric_utils.py:121-204: cv2.solvePnPRansac with
 * EPnP, iterationsCount 10000, reprojectionError `pnp_reprojection_error`, per frame on the CPU
 * after a D2H copy; callers: compute_query_pose_errors metric_utils.py:207-292, demo.py:132)
 * ---------------------------------------------------------------------------------------- */

/* Batched RANSAC-PnP, one CTA per image, consuming the matcher's output lists in place.
 *   pts3d fp32 [m][3] (mkpts_3d_db), pts2d fp32 [m][2] (mkpts_query_f), m_bids int64 [m] ascending
 *   (the matches of image b are the run m_bids == b); intrinsics fp32 [batch][3][3];
 *   scale: point-cloud rescale (3D points are multiplied by it, t is divided by it: metric_utils.py:179,193)
 *   reproj_thr: inlier threshold in pixels; hypotheses: P3P minimal samples per image;
 *   seed: RNG seed (counter based: results are reproducible and independent of scheduling);
 *   refine_rounds: local-optimisation rounds (Gauss-Newton on the inliers + inlier re-selection)
 * Outputs: poses fp32 [batch][3][4] = [R | t] (identity when the image fails), n_inliers int32
 * [batch], inlier_mask uint8 [m], status int32 [batch] (1 = pose found from >= 4 inliers). */
int opp_pnp_ransac(const float* pts3d, const float* pts2d, const long long* m_bids, int m,
                   const float* intrinsics, int batch, float scale, float reproj_thr, int hypotheses,
                   unsigned seed, int refine_rounds, float* poses, int* n_inliers,
                   unsigned char* inlier_mask, int* status, opp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* OPP_B200_H_ */
