/* A C host for libhvit_sm100.so: no Python, no torch - only the C ABI of include/hvit.h and the CUDA runtime.
 *
 *   enhance_host BLOB
 *
 * BLOB (written by tests/test_gpu_model.py::test_c_host_enhances_golden_clip) holds a model configuration, the
 * reference state_dict as raw fp32 tensors in registration order, a noisy clip and the waveform the REFERENCE's own
 * AudioEnhancer produced for it (tests/golden).  The program packs the weights (hvit_pack_weights), builds a plan,
 * runs hvit_enhance and prints the max relative waveform error against the reference.
 *
 * Build: gcc -std=c99 -I include -I /usr/local/cuda/include tests/host_c/enhance_host.c -o tests/host_c/enhance_host \
 *            -L <dir of libhvit_sm100.so> -lhvit_sm100 -L /usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,<dir>
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hvit.h"

#define CHECK_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)
#define CHECK_HVIT(x) do { int r_ = (x); if (r_ != HVIT_OK) { fprintf(stderr, "hvit error %d (%s) at %s:%d\n", r_, hvit_last_error(), __FILE__, __LINE__); return 3; } } while (0)

static FILE* g_f;

/* next tensor of the blob -> device memory */
static const float* take(size_t n) {
  float* h = (float*)malloc(n * sizeof(float));
  float* d = NULL;
  if (h == NULL || fread(h, sizeof(float), n, g_f) != n) { fprintf(stderr, "short blob\n"); exit(4); }
  if (cudaMalloc((void**)&d, n * sizeof(float)) != cudaSuccess || cudaMemcpy(d, h, n * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
    fprintf(stderr, "device upload failed\n");
    exit(4);
  }
  free(h);
  return d;
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s BLOB\n", argv[0]); return 1; }
  g_f = fopen(argv[1], "rb");
  if (g_f == NULL) { perror(argv[1]); return 1; }
  /* header: n_enc, enc_channels[n_enc], enc_pool[n_enc], embed_dim, num_heads, num_layers, mlp_hidden, patch,
   *         n_dec, dec_channels[n_dec], dec_up[n_dec], use_skip, precision, pos_len, n_samples */
  int hdr[64], nh = 0;
  if (fread(&nh, sizeof(int), 1, g_f) != 1 || nh > 64 || fread(hdr, sizeof(int), nh, g_f) != (size_t)nh) return 1;
  hvit_model_cfg cfg;
  memset(&cfg, 0, sizeof(cfg));
  int k = 0, i, l;
  cfg.n_enc = hdr[k++];
  for (i = 0; i < cfg.n_enc; ++i) cfg.enc_channels[i] = hdr[k++];
  for (i = 0; i < cfg.n_enc; ++i) cfg.enc_pool[i] = hdr[k++];
  cfg.embed_dim = hdr[k++]; cfg.num_heads = hdr[k++]; cfg.num_layers = hdr[k++]; cfg.mlp_hidden = hdr[k++]; cfg.patch_size = hdr[k++];
  cfg.n_dec = hdr[k++];
  for (i = 0; i < cfg.n_dec; ++i) cfg.dec_channels[i] = hdr[k++];
  for (i = 0; i < cfg.n_dec; ++i) cfg.dec_up[i] = hdr[k++];
  cfg.use_skip = hdr[k++]; cfg.precision = hdr[k++];
  const int pos_len = hdr[k++], n = hdr[k++];
  cfg.ln_eps = 1e-5f;
  if (hvit_device_ok() != 1) { fprintf(stderr, "no sm_100 device: %s\n", hvit_last_error()); return 5; }

  /* reference state_dict, registration order */
  hvit_ref_weights ref;
  memset(&ref, 0, sizeof(ref));
  const size_t D = (size_t)cfg.embed_dim, Hd = (size_t)cfg.mlp_hidden, P = (size_t)cfg.patch_size;
  size_t cin = 1;
  for (i = 0; i < cfg.n_enc; ++i) {
    const size_t c = (size_t)cfg.enc_channels[i];
    ref.enc_conv_w[i] = take(c * cin * 9);
    ref.enc_bn_w[i] = take(c); ref.enc_bn_b[i] = take(c); ref.enc_bn_mean[i] = take(c); ref.enc_bn_var[i] = take(c);
    cin = c;
  }
  const size_t clast = cin;
  ref.patch_w = take(D * clast * P * P);
  ref.patch_b = take(D);
  ref.pos_embed = take((size_t)pos_len * D);
  ref.pos_len = pos_len;
  for (l = 0; l < cfg.num_layers; ++l) {
    ref.ln1_w[l] = take(D); ref.ln1_b[l] = take(D); ref.ln2_w[l] = take(D); ref.ln2_b[l] = take(D);
    ref.qkv_w[l] = take(3 * D * D); ref.qkv_b[l] = take(3 * D);
    ref.proj_w[l] = take(D * D); ref.proj_b[l] = take(D);
    ref.fc1_w[l] = take(Hd * D); ref.fc1_b[l] = take(Hd);
    ref.fc2_w[l] = take(D * Hd); ref.fc2_b[l] = take(D);
  }
  ref.lnf_w = take(D); ref.lnf_b = take(D);
  ref.tofm_w = take(clast * D); ref.tofm_b = take(clast);
  for (i = 0; i < cfg.n_dec; ++i) {
    const int final_blk = i == cfg.n_dec - 1;
    const size_t c = (size_t)cfg.dec_channels[i];
    size_t ic = (size_t)(i == 0 ? cfg.dec_channels[0] : cfg.dec_channels[i - 1]);
    if (cfg.use_skip && !final_blk) ic += c;
    ref.dec_conv_w[i] = take(c * ic * 9);
    if (!final_blk) { ref.dec_bn_w[i] = take(c); ref.dec_bn_b[i] = take(c); ref.dec_bn_mean[i] = take(c); ref.dec_bn_var[i] = take(c); }
  }
  if (cfg.use_skip)
    for (i = 0; i < cfg.n_dec - 1 && i < cfg.n_enc; ++i) {
      ref.skip_w[i] = take((size_t)cfg.dec_channels[i] * (size_t)cfg.enc_channels[cfg.n_enc - 1 - i]);
      ref.skip_b[i] = take((size_t)cfg.dec_channels[i]);
    }
  float* noisy = (float*)malloc(sizeof(float) * (size_t)n);
  float* expect = (float*)malloc(sizeof(float) * (size_t)n);
  float* got = (float*)malloc(sizeof(float) * (size_t)n);
  if (fread(noisy, sizeof(float), (size_t)n, g_f) != (size_t)n || fread(expect, sizeof(float), (size_t)n, g_f) != (size_t)n) return 1;
  fclose(g_f);

  cudaStream_t stream;
  CHECK_CUDA(cudaStreamCreate(&stream));
  /* pack */
  const size_t pbytes = hvit_packed_weights_bytes(&cfg, pos_len);
  if (pbytes == 0) { fprintf(stderr, "packed size: %s\n", hvit_last_error()); return 3; }
  void* packed = NULL;
  CHECK_CUDA(cudaMalloc(&packed, pbytes));
  hvit_weights w;
  CHECK_HVIT(hvit_pack_weights(&cfg, &ref, packed, pbytes, &w, stream));
  /* plan */
  const int T = 1 + n / 128;
  const size_t wbytes = hvit_workspace_bytes(&cfg, 1, 257, T, n);
  if (wbytes == 0) { fprintf(stderr, "workspace size: %s\n", hvit_last_error()); return 3; }
  void* ws = NULL;
  CHECK_CUDA(cudaMalloc(&ws, wbytes));   /* cudaMalloc is 256-byte aligned at least; the plan wants 1024 */
  if (((size_t)ws & 1023) != 0) { fprintf(stderr, "workspace not 1024-byte aligned\n"); return 2; }
  hvit_plan* plan = NULL;
  CHECK_HVIT(hvit_plan_create(&cfg, &w, 1, 257, T, n, ws, wbytes, stream, &plan));
  /* enhance */
  float *d_in = NULL, *d_out = NULL;
  CHECK_CUDA(cudaMalloc((void**)&d_in, sizeof(float) * (size_t)n));
  CHECK_CUDA(cudaMalloc((void**)&d_out, sizeof(float) * (size_t)n));
  CHECK_CUDA(cudaMemcpyAsync(d_in, noisy, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, stream));
  CHECK_HVIT(hvit_enhance(plan, d_in, d_out, 1, stream));
  CHECK_CUDA(cudaMemcpyAsync(got, d_out, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, stream));
  CHECK_CUDA(cudaStreamSynchronize(stream));
  double dmax = 0.0, rmax = 0.0;
  for (i = 0; i < n; ++i) {
    const double d = fabs((double)got[i] - (double)expect[i]), r = fabs((double)expect[i]);
    if (!(d <= dmax)) dmax = d;   /* (a NaN makes the comparison false and propagates) */
    if (r > rmax) rmax = r;
  }
  printf("launches=%d packed_bytes=%zu workspace_bytes=%zu max_rel=%.6e\n", hvit_plan_launch_count(plan, 1), pbytes, wbytes,
         dmax / (rmax > 0 ? rmax : 1.0));
  hvit_plan_destroy(plan);
  return 0;
}
