// store_issue.cu -- developer microbenchmark: issue cost of 128-byte STG instructions (rows 1 KB apart) from
// 4 warps per SM, stamped with clock64 every 8 stores.  Variants: fresh region vs. region written before,
// 1 KB vs. 512 B row pitch, every SM vs. a single SM.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k(float* out, long long* stamps, int row_floats, int rows_per_warp, size_t cta_stride) {
  float* base = out + (size_t)blockIdx.x * cta_stride;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* w = base + (size_t)warp * rows_per_warp * row_floats + lane;
  long long* st = stamps + ((size_t)blockIdx.x * 4 + warp) * 64;
  int slot = 0;
  for (int r = 0; r < rows_per_warp; r += 8) {
    if (lane == 0) st[slot++] = clock64();
#pragma unroll
    for (int u = 0; u < 8; ++u) w[(size_t)(r + u) * row_floats] = (float)(r + u);
  }
  if (lane == 0) st[slot++] = clock64();
  __threadfence();
  if (lane == 0) st[slot++] = clock64();
}

int main() {
  float* buf; cudaMalloc(&buf, (size_t)1 << 30);
  long long* st; cudaMalloc(&st, 148 * 4 * 64 * 8);
  long long h[64];
  struct Cfg { const char* name; int ctas; int row_floats; int rows; bool warm; } cfgs[] = {
    {"148 CTAs, 1 KB pitch, 256 rows/warp, fresh", 148, 256, 256, false},
    {"148 CTAs, 1 KB pitch, 256 rows/warp, warm ", 148, 256, 256, true},
    {"  1 CTA , 1 KB pitch, 256 rows/warp, fresh", 1, 256, 256, false},
    {"148 CTAs, 128 B pitch (contiguous), fresh ", 148, 32, 256, false},
    {"148 CTAs, 4 KB pitch, 256 rows/warp, fresh", 148, 1024, 256, false},
  };
  size_t off = 0;
  for (auto& c : cfgs) {
    const size_t cta_stride = (size_t)4 * c.rows * c.row_floats;
    float* region = buf + off; off += (size_t)148 * cta_stride;
    if (c.warm) { k<<<c.ctas, 128>>>(region, st, c.row_floats, c.rows, cta_stride); cudaDeviceSynchronize(); }
    cudaMemset(st, 0, 148 * 4 * 64 * 8);
    k<<<c.ctas, 128>>>(region, st, c.row_floats, c.rows, cta_stride);
    cudaDeviceSynchronize();
    cudaMemcpy(h, st, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%s: cycles per 8 STG (warp 0 of CTA 0):", c.name);
    for (int i = 0; i < c.rows / 8; ++i) printf(" %lld", h[i + 1] - h[i]);
    printf("  | fence %lld\n", h[c.rows / 8 + 1] - h[c.rows / 8]);
  }
  printf("err: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
