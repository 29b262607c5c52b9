// embed_pool.cu -- K1 fused token-id gather + masked mean pool (fwd) and K2 its deterministic
// backward (dense embedding gradient).
//
// Reference: twotower/encoders.py:62,67,72 (mask = ids>0; W[ids]*mask; sum/(count+1e-9)) over
// nn.Embedding (twotower/embeddings.py:30,40); backward = ATen embedding_dense_backward
// reached from loss.backward() (twotower/train.py:138).
//
// fwd: one warp per batch row.  A row of the table is E floats; lanes are split into
//   32/TPT token-groups of TPT lanes (TPT = lanes needed to cover E/4 float4 chunks), so small
//   E (char tower, E=64 -> 16 lanes) gathers two tokens per warp instruction and large E
//   (E=300 -> 75 chunks) gives every lane up to 3 independent 128-bit loads per token.  The
//   [rows,L,E] tensor the reference materialises (twice) never exists.
// bwd, small tables (V <= 1024): build the pooling matrix P[r,v] = count(r,v) * inv_len[r]
//   (integer smem histogram -> exact, order-free) and compute dW = P^T * dPooled with the
//   fixed-order split-K GEMM.
// bwd, large tables: stable radix sort of (id, row) pairs, then a chunked segmented reduction
//   in sorted order: every vocabulary row has exactly one writer and a fixed summation order,
//   so dW is bitwise reproducible (atomics would not be).
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "sgemm.cuh"

namespace tt {

template <typename IdT>
__device__ __forceinline__ int64_t load_id(const IdT* p) { return (int64_t)__ldg(p); }

// The reference's nn.Embedding raises IndexError for an id outside [0, V) (embeddings.py:33-40).  A kernel cannot raise:
// it treats the token as padding (no out-of-bounds read) and records the offending id in a mapped host word that the
// host inspects at its next synchronisation point (tt_bad_token_id -> IndexError in the Python layer).
__device__ __forceinline__ void report_bad_id(long long* word, int64_t id) {
  if (word) *reinterpret_cast<volatile long long*>(word) = (long long)id * 2 + 1;
}

// ---------------------------------------------------------------------------------------
// K1 forward
// ---------------------------------------------------------------------------------------
template <typename IdT, int NACC, bool VEC>
__global__ void __launch_bounds__(256)
embed_pool_fwd_kernel(const IdT* __restrict__ ids, const float* __restrict__ table, int64_t rows,
                      int L, int64_t V, int E, int tpt, float* __restrict__ pooled,
                      float* __restrict__ inv_len, __nv_bfloat16* __restrict__ pooled_bf16,
                      __nv_bfloat16* __restrict__ pool_bf16, long long* __restrict__ bad_id) {
  extern __shared__ int fwd_hist[];                    // [warps][V] token histogram (only with pool_bf16)
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  int* hist = fwd_hist + (threadIdx.x >> 5) * (int)V;
  if (pool_bf16) {
    for (int v = lane; v < (int)V; v += 32) hist[v] = 0;
    __syncwarp();
  }
  const int groups = 32 / tpt;
  const int sub = lane % tpt, grp = lane / tpt;
  const IdT* rid = ids + row * L;

  if (VEC) {
    const int chunks = E >> 2;
    // column blocks of tpt*NACC float4 chunks (one pass for E <= 128*NACC)
    int count = 0;
    for (int cbase = 0; cbase < chunks; cbase += tpt * NACC) {
      float4 acc[NACC];
#pragma unroll
      for (int j = 0; j < NACC; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      int cnt = 0;
      // ids are read 32 tokens at a time with ONE coalesced load per lane and handed to the token
      // groups by shuffle, so the gather loads below do not wait on a per-token id load.
      constexpr int TOK = NACC == 1 ? 16 : (NACC == 2 ? 8 : 4);      // tokens in flight per group (16 LDG.128 per lane)
      int64_t next_id = (lane < L) ? load_id(rid + lane) : 0;
      for (int tb = 0; tb < L; tb += 32) {
        const int64_t my_id = next_id;
        if (tb + 32 < L) next_id = (tb + 32 + lane < L) ? load_id(rid + tb + 32 + lane) : 0;   // prefetch the next id block
        const int my_row = (my_id > 0 && my_id < V) ? (int)my_id : -1;       // -1 == masked token
        if (my_id < 0 || my_id >= V) report_bad_id(bad_id, my_id);           // nn.Embedding raises IndexError here
        cnt += __popc(__ballot_sync(0xffffffffu, my_row >= 0));
        if (pool_bf16 && cbase == 0 && my_row >= 0) atomicAdd(&hist[my_row], 1);   // integer counts: order-free, exact
        const int nb = min(32, L - tb);
        for (int tbase = 0; tbase < nb; tbase += TOK * groups) {     // warp-uniform trip count (shuffles inside)
          const int t0 = tbase + grp;
          int trow[TOK];
#pragma unroll
          for (int u = 0; u < TOK; ++u) {
            const int t = t0 + u * groups;
            const int r = __shfl_sync(0xffffffffu, my_row, t & 31);
            trow[u] = (t < nb) ? r : -1;
          }
          float4 v[TOK][NACC];
#pragma unroll
          for (int u = 0; u < TOK; ++u) {
            const float4* src = reinterpret_cast<const float4*>(table + (int64_t)(trow[u] < 0 ? 0 : trow[u]) * E);
#pragma unroll
            for (int j = 0; j < NACC; ++j) {
              const int c = cbase + sub + j * tpt;
              v[u][j] = (trow[u] >= 0 && c < chunks) ? __ldg(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
#pragma unroll
          for (int u = 0; u < TOK; ++u)
#pragma unroll
            for (int j = 0; j < NACC; ++j) {
              acc[j].x += v[u][j].x; acc[j].y += v[u][j].y; acc[j].z += v[u][j].z; acc[j].w += v[u][j].w;
            }
        }
      }
      // combine the token groups (lanes with equal `sub`), fixed xor-tree order
      for (int o = tpt; o < 32; o <<= 1) {
#pragma unroll
        for (int j = 0; j < NACC; ++j) {
          acc[j].x += __shfl_xor_sync(0xffffffffu, acc[j].x, o);
          acc[j].y += __shfl_xor_sync(0xffffffffu, acc[j].y, o);
          acc[j].z += __shfl_xor_sync(0xffffffffu, acc[j].z, o);
          acc[j].w += __shfl_xor_sync(0xffffffffu, acc[j].w, o);
        }
      }
      count = cnt;
      const float denom = (float)cnt + 1e-9f;           // encoders.py:72
      if (grp == 0) {
#pragma unroll
        for (int j = 0; j < NACC; ++j) {
          int c = cbase + sub + j * tpt;
          if (c < chunks) {
            float4 o = make_float4(acc[j].x / denom, acc[j].y / denom, acc[j].z / denom, acc[j].w / denom);
            *reinterpret_cast<float4*>(pooled + row * E + 4 * c) = o;
            if (pooled_bf16) {
              uint2 pk = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
              *reinterpret_cast<uint2*>(pooled_bf16 + row * E + 4 * c) = pk;
            }
          }
        }
      }
    }
    if (lane == 0 && inv_len) inv_len[row] = 1.0f / ((float)count + 1e-9f);
    if (pool_bf16) {                                     // P[row, v] = count(row, v) / len(row)  (the pooling matrix, bf16)
      __syncwarp();
      const float il = 1.0f / ((float)count + 1e-9f);
      for (int v = lane; v < (int)V; v += 32) pool_bf16[row * V + v] = __float2bfloat16((float)hist[v] * il);
    }
  } else {
    // scalar path (E not a multiple of 4 or unaligned table): lane-strided columns
    int cnt = 0;
    for (int t = 0; t < L; ++t) {
      int64_t id = load_id(rid + t);
      cnt += (id > 0 && id < V);
    }
    const float denom = (float)cnt + 1e-9f;
    for (int e = lane; e < E; e += 32) {
      float s = 0.f;
      for (int t = 0; t < L; ++t) {
        int64_t id = load_id(rid + t);
        if (id > 0 && id < V) s += __ldg(table + id * E + e);
      }
      float o = s / denom;
      pooled[row * E + e] = o;
      if (pooled_bf16) pooled_bf16[row * E + e] = __float2bfloat16(o);
    }
    if (lane == 0 && inv_len) inv_len[row] = 1.0f / denom;
    if (pool_bf16) {
      for (int t = lane; t < L; t += 32) {
        const int64_t id = load_id(rid + t);
        if (id > 0 && id < V) atomicAdd(&hist[id], 1);
      }
      __syncwarp();
      for (int v = lane; v < (int)V; v += 32) pool_bf16[row * V + v] = __float2bfloat16((float)hist[v] / denom);
    }
  }
}

// plain gather (API compatibility with LookupEmbedding.forward -> [B,L,E])
template <typename IdT>
__global__ void embed_gather_kernel(const IdT* __restrict__ ids, const float* __restrict__ table,
                                    int64_t n_tokens, int64_t V, int E, float* __restrict__ out, long long* __restrict__ bad_id) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t t = warp; t < n_tokens; t += nwarps) {
    int64_t id = load_id(ids + t);
    const bool ok = id >= 0 && id < V;
    if (!ok && lane == 0) report_bad_id(bad_id, id);
    for (int e = lane; e < E; e += 32) out[t * E + e] = ok ? __ldg(table + id * E + e) : 0.f;
  }
}

// ---------------------------------------------------------------------------------------
// K2 backward, small tables: pooling matrix + GEMM
// ---------------------------------------------------------------------------------------
template <typename IdT>
__global__ void __launch_bounds__(256)
pool_matrix_kernel(const IdT* __restrict__ ids, const float* __restrict__ inv_len, int64_t rows,
                   int L, int V, float* __restrict__ P) {
  extern __shared__ int hist[];                       // [warps][V]
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int* h = hist + w * V;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + w;
  if (row >= rows) return;
  for (int v = lane; v < V; v += 32) h[v] = 0;
  __syncwarp();
  for (int t = lane; t < L; t += 32) {
    int64_t id = load_id(ids + row * L + t);
    if (id > 0 && id < V) atomicAdd(&h[(int)id], 1);   // integer: exact, order-independent
  }
  __syncwarp();
  const float il = inv_len[row];
  for (int v = lane; v < V; v += 32) P[row * V + v] = (float)h[v] * il;
}

// Forward without the gather: only the pooling matrix P (bf16) and 1/len.  Used when the consumer forms x = P table on
// the tensor cores itself (tt_mlp_fwd with embed): the [rows,E] pooled tensor then never exists.
template <typename IdT>
__global__ void __launch_bounds__(256)
pool_only_kernel(const IdT* __restrict__ ids, int64_t rows, int L, int V, float* __restrict__ inv_len,
                 __nv_bfloat16* __restrict__ P, long long* __restrict__ bad_id) {
  extern __shared__ int hist[];                       // [warps][V]
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int* h = hist + w * V;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + w;
  if (row >= rows) return;
  for (int v = lane; v < V; v += 32) h[v] = 0;
  __syncwarp();
  int cnt = 0;
  for (int t0 = 0; t0 < L; t0 += 32) {
    const int t = t0 + lane;
    const int64_t id = t < L ? load_id(ids + row * L + t) : 0;
    const bool ok = id > 0 && id < V;
    if (id < 0 || id >= V) report_bad_id(bad_id, id);
    if (ok) atomicAdd(&h[(int)id], 1);                // integer: exact, order-independent
    cnt += __popc(__ballot_sync(0xffffffffu, ok));
  }
  __syncwarp();
  const float il = 1.0f / ((float)cnt + 1e-9f);       // encoders.py:72
  if (lane == 0 && inv_len) inv_len[row] = il;
  for (int v = lane * 2; v < V; v += 64) {             // V is even on this path (V % 8 == 0)
    *reinterpret_cast<uint32_t*>(P + row * V + v) = pack_bf16x2((float)h[v] * il, (float)h[v + 1] * il);
  }
}

// ---------------------------------------------------------------------------------------
// K2 backward, large tables: sort + ordered segmented reduction
// ---------------------------------------------------------------------------------------
template <typename IdT>
__global__ void make_keys_kernel(const IdT* __restrict__ ids, int64_t n_tokens, int L, int64_t V,
                                 uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n_tokens;
       t += (int64_t)gridDim.x * blockDim.x) {
    int64_t id = load_id(ids + t);
    keys[t] = (id > 0 && id < V) ? (uint32_t)id : (uint32_t)V;   // masked tokens sort last
    vals[t] = (uint32_t)(t / L);                                 // owning batch row
  }
}

constexpr int kSegChunk = 64;     // sorted tokens per warp

// One warp per chunk of kSegChunk sorted tokens.  Segments that START inside the chunk are
// written straight to dW (sole writer).  The leading run that continues a segment begun in an
// earlier chunk goes to carry[chunk]; seg_fixup adds carries in chunk order.
// The chunk's (id, row, 1/len) triples are read once, coalesced, into registers (2 per lane) and handed out by
// shuffle; the gradient row of token t+1 is already in flight while token t is accumulated, and a lane owns NC float4
// column chunks of the whole row (E <= 128 * NC), so the sorted token list is walked ONCE.  With per-token dependent
// loads (id -> row -> 1/len -> gradient row) this kernel ran at the latency of ~200 serial round trips per warp.
template <int NC>
__global__ void __launch_bounds__(256)
seg_reduce_vec_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                      const float* __restrict__ inv_len, const float* __restrict__ g, int64_t n_tokens,
                      int64_t V, int E, float* __restrict__ dW, float* __restrict__ carry) {
  const int lane = threadIdx.x & 31;
  const int64_t chunk = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t p0 = chunk * kSegChunk;
  if (p0 >= n_tokens) return;
  const int cnt = (int)((p0 + kSegChunk < n_tokens ? p0 + kSegChunk : n_tokens) - p0);
  const int nchunks = E >> 2;                              // float4 chunks per row (E % 4 == 0 on this path)
  uint32_t k2[2], r2[2];
  float s2[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int t = lane + 32 * u;
    k2[u] = t < cnt ? __ldg(keys + p0 + t) : 0xffffffffu;
    r2[u] = t < cnt ? __ldg(vals + p0 + t) : 0u;
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) s2[u] = (lane + 32 * u < cnt && k2[u] < (uint32_t)V) ? __ldg(inv_len + r2[u]) : 0.f;
  uint32_t prev = (p0 > 0 && lane == 0) ? __ldg(keys + p0 - 1) : 0xffffffffu;
  prev = __shfl_sync(0xffffffffu, prev, 0);
  auto tok = [&](int t, uint32_t& key, uint32_t& row, float& sc) {
    const int src = t & 31;
    const uint32_t ka = __shfl_sync(0xffffffffu, k2[0], src), kb = __shfl_sync(0xffffffffu, k2[1], src);
    const uint32_t ra = __shfl_sync(0xffffffffu, r2[0], src), rb = __shfl_sync(0xffffffffu, r2[1], src);
    const float sa = __shfl_sync(0xffffffffu, s2[0], src), sb = __shfl_sync(0xffffffffu, s2[1], src);
    key = t < 32 ? ka : kb; row = t < 32 ? ra : rb; sc = t < 32 ? sa : sb;
  };
  auto load_row = [&](uint32_t key, uint32_t row, float4 (&v)[NC]) {
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const int c = lane + 32 * j;
      v[j] = (key < (uint32_t)V && c < nchunks) ? __ldg(reinterpret_cast<const float4*>(g + (int64_t)row * E) + c)
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  float4 acc[NC], cur_v[NC], nxt_v[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  uint32_t key, row, nkey = 0xffffffffu, nrow = 0u;
  float sc, nsc = 0.f;
  tok(0, key, row, sc);
  load_row(key, row, cur_v);
  uint32_t cur = key;
  bool continued = (p0 > 0) && (prev == cur);
  auto flush = [&]() {
    if (cur < (uint32_t)V) {
      float* dst = continued ? carry + chunk * E : dW + (int64_t)cur * E;
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        const int c = lane + 32 * j;
        if (c < nchunks) reinterpret_cast<float4*>(dst)[c] = acc[j];
      }
    }
#pragma unroll
    for (int j = 0; j < NC; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  };
  for (int t = 0; t < cnt; ++t) {
    if (t + 1 < cnt) { tok(t + 1, nkey, nrow, nsc); load_row(nkey, nrow, nxt_v); }   // next gradient row in flight
    if (key != cur) { flush(); cur = key; continued = false; }
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      acc[j].x = fmaf(cur_v[j].x, sc, acc[j].x); acc[j].y = fmaf(cur_v[j].y, sc, acc[j].y);
      acc[j].z = fmaf(cur_v[j].z, sc, acc[j].z); acc[j].w = fmaf(cur_v[j].w, sc, acc[j].w);
    }
    key = nkey; row = nrow; sc = nsc;
#pragma unroll
    for (int j = 0; j < NC; ++j) cur_v[j] = nxt_v[j];
  }
  flush();
}

// generic E (not a multiple of 4, or E > 512): scalar columns, 128 per pass
__global__ void __launch_bounds__(256)
seg_reduce_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                  const float* __restrict__ inv_len, const float* __restrict__ g, int64_t n_tokens,
                  int64_t V, int E, float* __restrict__ dW, float* __restrict__ carry) {
  const int lane = threadIdx.x & 31;
  const int64_t chunk = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t p0 = chunk * kSegChunk;
  if (p0 >= n_tokens) return;
  const int64_t p1 = (p0 + kSegChunk < n_tokens) ? p0 + kSegChunk : n_tokens;
  for (int ebase = 0; ebase < E; ebase += 128) {          // 4 columns per lane per pass
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    uint32_t cur = keys[p0];
    bool continued = (p0 > 0) && (keys[p0 - 1] == cur);
    for (int64_t p = p0; p <= p1; ++p) {
      uint32_t k = (p < p1) ? keys[p] : 0xffffffffu;
      if (k != cur) {                                     // flush finished run
        if (cur < (uint32_t)V) {
          float* dst = continued ? carry + chunk * E : dW + (int64_t)cur * E;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            int e = ebase + lane + 32 * j;
            if (e < E) dst[e] = acc[j];
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = 0.f;
        cur = k;
        continued = false;
        if (p == p1) break;
      }
      const uint32_t r = vals[p];
      const float s = inv_len[r];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int e = ebase + lane + 32 * j;
        if (e < E) acc[j] = fmaf(__ldg(g + (int64_t)r * E + e), s, acc[j]);
      }
    }
  }
}

// One warp per chunk whose LAST run continues into later chunks and STARTED in this chunk (or at token 0): it owns the
// fix-up of that id.  The end of the run is found by a binary search over the sorted keys; the carries of the chunks
// the run passes through are then added in chunk order (fixed order -> bitwise reproducible), every lane streaming its
// own columns with independent loads.  Zipf-distributed ids put ~7 % of all tokens on the hottest row: that run spans
// hundreds of chunks, which a chunk-by-chunk read-modify-write of dW walked at one memory round trip per chunk.
__global__ void __launch_bounds__(256)
seg_fixup_kernel(const uint32_t* __restrict__ keys, int64_t n_tokens, int64_t V, int E,
                 float* __restrict__ dW, const float* __restrict__ carry) {
  const int lane = threadIdx.x & 31;
  const int64_t chunk = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t p0 = chunk * kSegChunk;
  if (p0 >= n_tokens) return;
  const int64_t p1 = (p0 + kSegChunk < n_tokens) ? p0 + kSegChunk : n_tokens;
  if (p1 >= n_tokens) return;                             // nothing after this chunk
  const uint32_t key = keys[p1 - 1];
  if (key >= (uint32_t)V || keys[p1] != key) return;      // last run ends here
  // the run must have started in this chunk, otherwise an earlier chunk owns the fix-up
  if (keys[p0] == key && p0 > 0 && keys[p0 - 1] == key) return;
  int64_t lo = p1, hi = n_tokens;                         // first position in [p1, n) whose key differs (keys are sorted)
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (keys[mid] == key) lo = mid + 1; else hi = mid;
  }
  const int64_t c_last = (lo - 1) / kSegChunk;            // chunk holding the last token of the run
  float* dst = dW + (int64_t)key * E;
  // blockIdx.y selects a group of 128 columns (4 per lane), so the hottest rows' fix-ups are spread over E/128 warps;
  // 8 chunks of carries are in flight per step and still added in chunk order
  const int e0 = blockIdx.y * 128;
  float acc[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { const int e = e0 + lane + 32 * j; acc[j] = e < E ? dst[e] : 0.f; }
  for (int64_t c = chunk + 1; c <= c_last; c += 8) {
    float v[8][4];
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int e = e0 + lane + 32 * j;
        v[u][j] = (c + u <= c_last && e < E) ? __ldg(carry + (c + u) * E + e) : 0.f;
      }
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] += v[u][j];
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) { const int e = e0 + lane + 32 * j; if (e < E) dst[e] = acc[j]; }
}

constexpr int64_t kSmallVocab = 1024;

struct BwdPlan {
  bool small;
  int splits;
  size_t p_bytes, partial_bytes;                 // small path
  size_t keys_bytes, carry_bytes, cub_bytes;     // sort path
  size_t total;
};

static BwdPlan plan_bwd(int64_t rows, int L, int64_t V, int E) {
  BwdPlan p{};
  p.small = (V <= kSmallVocab);
  if (p.small) {
    p.splits = sgemm_pick_splits((int)V, E, (int)rows);
    p.p_bytes = align_up((size_t)rows * V * sizeof(float));
    p.partial_bytes = align_up(sgemm_partial_bytes((int)V, E, p.splits));
    p.total = p.p_bytes + p.partial_bytes;
  } else {
    const int64_t n = rows * L;
    p.keys_bytes = align_up((size_t)n * sizeof(uint32_t));
    p.carry_bytes = align_up((size_t)ceil_div(n, kSegChunk) * E * sizeof(float));
    size_t cub = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n, 0, 32);
    p.cub_bytes = align_up(cub);
    p.total = 4 * p.keys_bytes + p.carry_bytes + p.cub_bytes;
  }
  p.total += 256;
  return p;
}

template <typename IdT>
static int embed_pool_fwd_t(const IdT* ids, const float* table, int64_t rows, int L, int64_t V, int E,
                            float* pooled, float* inv_len, __nv_bfloat16* pooled_bf16, __nv_bfloat16* pool_bf16,
                            cudaStream_t s) {
  const int warps = 8;
  const unsigned grid = (unsigned)ceil_div(rows, warps);
  if (pool_bf16 && V > kSmallVocab) { set_error("embed_pool_fwd: the pooling-matrix output needs V <= %d", kSmallVocab); return TT_ERR_UNSUPPORTED; }
  const size_t hsm = pool_bf16 ? (size_t)warps * V * sizeof(int) : 0;
  if (pooled == nullptr) {                             // histogram only: the consumer multiplies P by the table itself
    if (!pool_bf16 || V % 8 != 0) { set_error("embed_pool_fwd: pooled == NULL needs pool_bf16 and V %% 8 == 0"); return TT_ERR_INVALID; }
    TT_CUDA(launch_kernel(pool_only_kernel<IdT>, dim3(grid), dim3(warps * 32), hsm, s, true, ids, rows, L, (int)V, inv_len, pool_bf16, bad_id_word()));
    TT_LAUNCH_CHECK("pool_only_kernel");
    return TT_OK;
  }
  const bool vec = (E % 4 == 0) && ((reinterpret_cast<uintptr_t>(table) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(pooled) & 15) == 0) &&
                   (pooled_bf16 == nullptr || (reinterpret_cast<uintptr_t>(pooled_bf16) & 7) == 0);
  if (!vec) {
    TT_CUDA(launch_kernel(embed_pool_fwd_kernel<IdT, 1, false>, dim3(grid), dim3(warps * 32), hsm, s, true, ids, table, rows, L, V, E, 32,
                          pooled, inv_len, pooled_bf16, pool_bf16, bad_id_word()));
  } else {
    const int chunks = E / 4;
    int tpt = 1;
    while (tpt < chunks && tpt < 32) tpt <<= 1;
    const int per_lane = (int)ceil_div(chunks, tpt);
    if (per_lane <= 1)
      TT_CUDA(launch_kernel(embed_pool_fwd_kernel<IdT, 1, true>, dim3(grid), dim3(warps * 32), hsm, s, true, ids, table, rows, L, V, E, tpt, pooled, inv_len, pooled_bf16, pool_bf16, bad_id_word()));
    else if (per_lane == 2)
      TT_CUDA(launch_kernel(embed_pool_fwd_kernel<IdT, 2, true>, dim3(grid), dim3(warps * 32), hsm, s, true, ids, table, rows, L, V, E, tpt, pooled, inv_len, pooled_bf16, pool_bf16, bad_id_word()));
    else
      TT_CUDA(launch_kernel(embed_pool_fwd_kernel<IdT, 3, true>, dim3(grid), dim3(warps * 32), hsm, s, true, ids, table, rows, L, V, E, tpt, pooled, inv_len, pooled_bf16, pool_bf16, bad_id_word()));
  }
  TT_LAUNCH_CHECK("embed_pool_fwd_kernel");
  return TT_OK;
}

template <typename IdT>
static int embed_pool_bwd_t(const IdT* ids, const float* inv_len, const float* d_pooled, int64_t rows,
                            int L, int64_t V, int E, float* d_table, void* ws, size_t ws_bytes,
                            cudaStream_t s) {
  const BwdPlan plan = plan_bwd(rows, L, V, E);
  if (ws == nullptr || ws_bytes < plan.total) {
    set_error("embed_pool_bwd: workspace too small (%zu < %zu)", ws_bytes, plan.total);
    return TT_ERR_WORKSPACE;
  }
  Workspace w(ws, ws_bytes);
  if (plan.small) {
    float* P = w.take<float>((size_t)rows * V);
    float* partial = plan.splits > 1 ? w.take<float>((size_t)plan.splits * V * E) : nullptr;
    const int warps = 8;
    TT_CUDA(launch_kernel(pool_matrix_kernel<IdT>, dim3((unsigned)ceil_div(rows, warps)), dim3(warps * 32), warps * V * sizeof(int), s, true,
                          ids, inv_len, rows, L, (int)V, P));
    TT_LAUNCH_CHECK("pool_matrix_kernel");
    SgemmArgs a{};
    a.M = (int)V; a.N = E; a.K = (int)rows;
    a.A = P; a.lda = (int)V; a.transA = 1;           // P^T
    a.B = d_pooled; a.ldb = E; a.transB = 0;
    a.C = d_table; a.ldc = E;
    a.splits = plan.splits; a.partial = partial;
    return sgemm(a, s);                               // row 0 of P is all-zero -> dW[0] = 0
  }
  const int64_t n = rows * L;
  TT_CHECK_ARG(n < (int64_t)1 << 31, "embed_pool_bwd: too many tokens (%lld)", (long long)n);
  uint32_t* keys_in = w.take<uint32_t>(n);
  uint32_t* vals_in = w.take<uint32_t>(n);
  uint32_t* keys = w.take<uint32_t>(n);
  uint32_t* vals = w.take<uint32_t>(n);
  float* carry = w.take<float>((size_t)ceil_div(n, kSegChunk) * E);
  void* cub_ws = w.take<char>(plan.cub_bytes);
  size_t cub_bytes = plan.cub_bytes;
  TT_CUDA(cudaMemsetAsync(d_table, 0, (size_t)V * E * sizeof(float), s));
  make_keys_kernel<IdT><<<(unsigned)(ceil_div(n, 256) < 8 * kNumSMs ? ceil_div(n, 256) : 8 * kNumSMs), 256, 0, s>>>(
      ids, n, L, V, keys_in, vals_in);
  TT_LAUNCH_CHECK("make_keys_kernel");
  int end_bit = 1;
  while (((int64_t)1 << end_bit) <= V && end_bit < 32) ++end_bit;
  TT_CUDA(cub::DeviceRadixSort::SortPairs(cub_ws, cub_bytes, keys_in, keys, vals_in, vals, (int)n, 0,
                                          end_bit, s));      // stable -> fixed order
  count_launch(3);
  const int warps = 8;
  const unsigned grid = (unsigned)ceil_div(ceil_div(n, kSegChunk), warps);
  const bool vec = (E % 4 == 0) && E <= 512 && ((reinterpret_cast<uintptr_t>(d_pooled) | reinterpret_cast<uintptr_t>(d_table) |
                                                  reinterpret_cast<uintptr_t>(carry)) & 15) == 0;
  if (vec && E <= 128)      seg_reduce_vec_kernel<1><<<grid, warps * 32, 0, s>>>(keys, vals, inv_len, d_pooled, n, V, E, d_table, carry);
  else if (vec && E <= 256) seg_reduce_vec_kernel<2><<<grid, warps * 32, 0, s>>>(keys, vals, inv_len, d_pooled, n, V, E, d_table, carry);
  else if (vec && E <= 384) seg_reduce_vec_kernel<3><<<grid, warps * 32, 0, s>>>(keys, vals, inv_len, d_pooled, n, V, E, d_table, carry);
  else if (vec)             seg_reduce_vec_kernel<4><<<grid, warps * 32, 0, s>>>(keys, vals, inv_len, d_pooled, n, V, E, d_table, carry);
  else                      seg_reduce_kernel<<<grid, warps * 32, 0, s>>>(keys, vals, inv_len, d_pooled, n, V, E, d_table, carry);
  TT_LAUNCH_CHECK("seg_reduce_kernel");
  seg_fixup_kernel<<<dim3(grid, (unsigned)ceil_div(E, 128)), warps * 32, 0, s>>>(keys, n, V, E, d_table, carry);
  TT_LAUNCH_CHECK("seg_fixup_kernel");
  return TT_OK;
}

}  // namespace tt

extern "C" {

int tt_embed_gather(const void* ids, int id_bytes, const float* table, int64_t n_tokens, int64_t V,
                    int E, float* out, void* stream) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG(ids && table && out && n_tokens >= 0 && V > 0 && E > 0, "embed_gather: bad arguments");
  TT_CHECK_ARG(id_bytes == 4 || id_bytes == 8, "embed_gather: id_bytes must be 4 or 8");
  if (n_tokens == 0) return TT_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  unsigned grid = (unsigned)(tt::ceil_div(n_tokens, 8) < 16 * tt::kNumSMs ? tt::ceil_div(n_tokens, 8) : 16 * tt::kNumSMs);
  if (id_bytes == 8)
    tt::embed_gather_kernel<int64_t><<<grid, 256, 0, s>>>((const int64_t*)ids, table, n_tokens, V, E, out, tt::bad_id_word());
  else
    tt::embed_gather_kernel<int32_t><<<grid, 256, 0, s>>>((const int32_t*)ids, table, n_tokens, V, E, out, tt::bad_id_word());
  TT_LAUNCH_CHECK("embed_gather_kernel");
  return TT_OK;
}

int tt_embed_pool_fwd(const void* ids, int id_bytes, const float* table, int64_t rows, int L, int64_t V,
                      int E, float* pooled, float* inv_len, void* pooled_bf16, void* pool_bf16, void* stream) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG(ids && table && (pooled || pool_bf16) && rows >= 0 && L > 0 && V > 0 && E > 0, "embed_pool_fwd: bad arguments");
  TT_CHECK_ARG(id_bytes == 4 || id_bytes == 8, "embed_pool_fwd: id_bytes must be 4 or 8");
  if (rows == 0) return TT_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (id_bytes == 8)
    return tt::embed_pool_fwd_t<int64_t>((const int64_t*)ids, table, rows, L, V, E, pooled, inv_len,
                                         (__nv_bfloat16*)pooled_bf16, (__nv_bfloat16*)pool_bf16, s);
  return tt::embed_pool_fwd_t<int32_t>((const int32_t*)ids, table, rows, L, V, E, pooled, inv_len,
                                       (__nv_bfloat16*)pooled_bf16, (__nv_bfloat16*)pool_bf16, s);
}

size_t tt_embed_pool_bwd_workspace(int64_t rows, int L, int64_t V, int E) {
  if (rows <= 0 || L <= 0 || V <= 0 || E <= 0) return 256;
  return tt::plan_bwd(rows, L, V, E).total;
}

int tt_embed_pool_bwd(const void* ids, int id_bytes, const float* inv_len, const float* d_pooled,
                      int64_t rows, int L, int64_t V, int E, float* d_table, void* workspace,
                      size_t workspace_bytes, void* stream) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG(ids && inv_len && d_pooled && d_table && rows >= 0 && L > 0 && V > 0 && E > 0,
               "embed_pool_bwd: bad arguments");
  TT_CHECK_ARG(id_bytes == 4 || id_bytes == 8, "embed_pool_bwd: id_bytes must be 4 or 8");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (rows == 0) {
    TT_CUDA(cudaMemsetAsync(d_table, 0, (size_t)V * E * sizeof(float), s));
    return TT_OK;
  }
  if (id_bytes == 8)
    return tt::embed_pool_bwd_t<int64_t>((const int64_t*)ids, inv_len, d_pooled, rows, L, V, E, d_table,
                                         workspace, workspace_bytes, s);
  return tt::embed_pool_bwd_t<int32_t>((const int32_t*)ids, inv_len, d_pooled, rows, L, V, E, d_table,
                                       workspace, workspace_bytes, s);
}

}  // extern "C"
