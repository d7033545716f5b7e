// Fused CombinedLoss of the reference (src/losses/combined_loss.py:64-191; SURVEY.md section 8f "next-2"): the four terms
// and d(total)/d(pred) in HBM-bound passes over pred [B, T, Hs, Ws] fp32 (C = 1):
//   pass A  (per LR cell)   : area pooling s x s -> |mean - lr| (conservation, :64-74) and the sign of the residual
//   pass B  (per HR pixel)  : |dx| + |dy| (spatial gradient, :146-155), |dt| (temporal, :160-168), and the gradient of
//                             all three grid terms gathered from the pixel's own edges + its LR cell's sign
//                             (A and B run in ONE kernel, band by band, so pred comes from HBM once)
//   pass C  (per station obs): weighted L1 at the gauge pixels (:79-141), NaN observations skipped; its gradient is
//                             scattered (atomicAdd) onto dpred
// Means follow torch exactly: each term is divided by ITS OWN element count; sign(0) = 0 (abs backward).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace plc {

struct LossParams {
  int B, T, H, W;          // LR grid
  int s;                   // integer upsampling ratio: Hs = H*s, Ws = W*s
  int Hs, Ws;
  int n_st;                // stations
  int svals_has_batch;     // s_values is [B,T,N] (1) or [T,N] (0)
  float coord_scale;       // scale_factor used for the station coordinates (combined_loss.py:97)
  int weight_mode;         // 0 none, 1 log (1 + log1p), 2 sqrt (1 + sqrt), 3 stratified
  float l_point, l_cons, l_smooth, l_temp;
  const float* pred;       // [B,T,Hs,Ws]
  const float* lr;         // [B,T,H,W]
  const long long* coords; // [N,2] (row, col) on the LR grid
  const float* svals;      // [T,N] or [B,T,N], NaN = missing
  float* sums;             // [8]: 0 cons, 1 gx, 2 gy, 3 temporal, 4 point (weighted), 5 point count
  float* dpred;            // [B,T,Hs,Ws] or nullptr
  const float* grad_scale; // device scalar multiplied into dpred (upstream gradient), or nullptr
};

__device__ __forceinline__ float sgn(float v) { return (v > 0.f) - (v < 0.f); }
// c * sign(e), sign(0) = 0 (torch's abs backward)
__device__ __forceinline__ float sgn_scaled(float c, float e) { return e == 0.f ? 0.f : copysignf(c, e); }

__device__ __forceinline__ float block_sum(float v, float* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) red[w] = v;
  __syncthreads();
  float r = 0.f;
  if (w == 0) {
    r = l < (blockDim.x >> 5) ? red[l] : 0.f;
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  __syncthreads();
  return r;   // valid in warp 0
}

// passes A + B fused: one block per (image bt, band of m LR rows = R = m*s HR rows).
//   phase 1: one thread per LR cell of the band sums its s x s pixels (first touch: HBM), keeps sign(mean - lr) in smem
//   phase 2: one thread per 4 HR pixels (float4 when Ws % 4 == 0) re-reads the band (L1/L2-hot), gathers the gradient
//            of the smooth / temporal / conservation terms and stores dpred
// HBM traffic ~ 4 B read + 4 B written per HR pixel; the t+-1 frames and the band's halo rows are L2 hits.
// blockDim = (bx, by = 256 / bx), bx = power of two covering a row; thread (tx, ty) owns rows [ty*rpt, (ty+1)*rpt) of
// the band for its columns.  Dynamic smem: m * W floats.
template <int VEC>
__global__ void __launch_bounds__(256) loss_grid_kernel(const LossParams p, int m, int bands, int nblk, int rpt) {
  extern __shared__ float s_sign[];          // [m][W]
  __shared__ float red[8];
  float acons = 0.f, agx = 0.f, agy = 0.f, at = 0.f;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  // persistent blocks: one atomicAdd per term per BLOCK at the end, not per band (same-address atomics serialise)
  for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
  const int bt = blk / bands;
  const int band = blk - bt * bands;
  const int Y0 = band * m;                                   // first LR row of the band
  const int mrows = min(m, p.H - Y0);
  const size_t hw = static_cast<size_t>(p.Hs) * p.Ws;
  const float* img = p.pred + static_cast<size_t>(bt) * hw;
  const int t = bt % p.T;
  const float inv_ss = 1.f / (p.s * p.s);

  // ---- phase 1: conservation residual per LR cell (s == 1: a cell is a pixel, handled in phase 2)
  const float* lr = p.lr + (static_cast<size_t>(bt) * p.H + Y0) * p.W;
  if (p.s > 1) {
    for (int c = tid; c < mrows * p.W; c += 256) {
      const int yy = c / p.W, X = c - yy * p.W;
      const float* src = img + static_cast<size_t>(Y0 + yy) * p.s * p.Ws + static_cast<size_t>(X) * p.s;
      float sum = 0.f;
      if (VEC == 4 && (p.s & 3) == 0) {
        for (int dy = 0; dy < p.s; ++dy)
          for (int dx = 0; dx < p.s; dx += 4) {
            const float4 v = *reinterpret_cast<const float4*>(src + static_cast<size_t>(dy) * p.Ws + dx);
            sum += (v.x + v.y) + (v.z + v.w);
          }
      } else {
        for (int dy = 0; dy < p.s; ++dy)
          for (int dx = 0; dx < p.s; ++dx) sum += src[static_cast<size_t>(dy) * p.Ws + dx];
      }
      const float d = sum * inv_ss - lr[c];                  // F.interpolate(mode="area") for an integer ratio
      acons += fabsf(d);
      s_sign[c] = sgn(d);
    }
    __syncthreads();
  }

  // ---- phase 2: per-pixel terms + gradient
  const float n_bt = static_cast<float>(p.B) * p.T;
  const float up_g = p.grad_scale ? *p.grad_scale : 1.f;
  const float cgx = p.Ws > 1 ? up_g * p.l_smooth / (n_bt * p.Hs * (p.Ws - 1)) : 0.f;
  const float cgy = p.Hs > 1 ? up_g * p.l_smooth / (n_bt * (p.Hs - 1) * p.Ws) : 0.f;
  const float ct = p.T > 1 ? up_g * p.l_temp / (static_cast<float>(p.B) * (p.T - 1) * hw) : 0.f;
  const float cc = up_g * p.l_cons / (n_bt * p.H * p.W) * inv_ss;
  const int y0 = Y0 * p.s, rows = mrows * p.s;
  const bool has_nx = t + 1 < p.T, has_pv = t > 0;
  const bool one_cell = VEC == 1 || (p.s & 3) == 0;          // the thread's pixels share one LR cell
  // each thread walks a strip of `rpt` rows x VEC columns top to bottom, carrying the row above in registers:
  // every vertical edge inside a strip is evaluated once and feeds both of its pixels
  const int ya = y0 + threadIdx.y * rpt, yb = min(y0 + rows, ya + rpt);
  for (int x = threadIdx.x * VEC; x < p.Ws && ya < yb; x += blockDim.x * VEC) {
    const float* col = img + x;
    float cur[VEC], nxt[VEC], o[VEC], g[VEC], carry[VEC];
    auto load = [&](float* dst, const float* src) {
      if (VEC == 4) *reinterpret_cast<float4*>(dst) = *reinterpret_cast<const float4*>(src);
      else dst[0] = src[0];
    };
    load(cur, col + static_cast<size_t>(ya) * p.Ws);
#pragma unroll
    for (int j = 0; j < VEC; ++j) carry[j] = 0.f;
    if (ya > 0) {                                            // edge to the row above the strip (summed by its owner)
      load(o, col + static_cast<size_t>(ya - 1) * p.Ws);
#pragma unroll
      for (int j = 0; j < VEC; ++j) carry[j] = -sgn_scaled(cgy, o[j] - cur[j]);
    }
    const int cx = x / p.s;
    int cy = (ya - y0) / p.s, crem = (ya - y0) - cy * p.s;
    const float* row = col + static_cast<size_t>(ya) * p.Ws;
    for (int y = ya; y < yb; ++y, row += p.Ws) {
      const bool has_dn = y + 1 < p.Hs;
      if (has_dn) load(nxt, row + p.Ws);
#pragma unroll
      for (int j = 0; j < VEC; ++j) g[j] = carry[j];
      // horizontal edges: VEC + 1 of them, each evaluated once; the left one is summed by the thread on the left
      if (x > 0) g[0] -= sgn_scaled(cgx, row[-1] - cur[0]);
#pragma unroll
      for (int j = 0; j + 1 < VEC; ++j) {
        const float e = cur[j] - cur[j + 1];
        agx += fabsf(e);
        const float sg = sgn_scaled(cgx, e);
        g[j] += sg;
        g[j + 1] -= sg;
      }
      if (x + VEC < p.Ws) {
        const float e = cur[VEC - 1] - row[VEC];
        agx += fabsf(e);
        g[VEC - 1] += sgn_scaled(cgx, e);
      }
      if (has_dn) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const float e = cur[j] - nxt[j];
          agy += fabsf(e);
          const float sg = sgn_scaled(cgy, e);
          g[j] += sg;
          carry[j] = -sg;
        }
      }
      if (has_nx) {
        load(o, row + hw);
#pragma unroll
        for (int j = 0; j < VEC; ++j) { const float e = cur[j] - o[j]; at += fabsf(e); g[j] += sgn_scaled(ct, e); }
      }
      if (has_pv) {
        load(o, row - hw);
#pragma unroll
        for (int j = 0; j < VEC; ++j) g[j] -= sgn_scaled(ct, o[j] - cur[j]);
      }
      if (p.s == 1) {                                        // conservation on the pixel itself
        load(o, lr + static_cast<size_t>(y - y0) * p.W + x);
#pragma unroll
        for (int j = 0; j < VEC; ++j) { const float d = cur[j] - o[j]; acons += fabsf(d); g[j] += sgn_scaled(cc, d); }
      } else if (one_cell) {
        const float cs = cc * s_sign[cy * p.W + cx];
#pragma unroll
        for (int j = 0; j < VEC; ++j) g[j] += cs;
      } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) g[j] += cc * s_sign[cy * p.W + (x + j) / p.s];
      }
      if (p.dpred) {
        float* dst = p.dpred + (row - p.pred);
        if (VEC == 4) *reinterpret_cast<float4*>(dst) = *reinterpret_cast<float4*>(g);
        else dst[0] = g[0];
      }
#pragma unroll
      for (int j = 0; j < VEC; ++j) cur[j] = nxt[j];
      if (++crem == p.s) { crem = 0; ++cy; }
    }
  }
  __syncthreads();                                           // s_sign is rewritten by the next band
  }
  auto bsum = [&](float val) {
    for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
    const int w = tid >> 5, l = tid & 31;
    if (l == 0) red[w] = val;
    __syncthreads();
    float r = 0.f;
    if (w == 0) {
      r = l < 8 ? red[l] : 0.f;
      for (int o = 4; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    __syncthreads();
    return r;
  };
  float r = bsum(acons);
  if (tid == 0) atomicAdd(p.sums + 0, r);
  r = bsum(agx);
  if (tid == 0) atomicAdd(p.sums + 1, r);
  r = bsum(agy);
  if (tid == 0) atomicAdd(p.sums + 2, r);
  r = bsum(at);
  if (tid == 0) atomicAdd(p.sums + 3, r);
}

__device__ __forceinline__ bool station_pixel(const LossParams& p, int st, int& row, int& col) {
  // ((coords + 0.5) * scale - 0.5).long()  -- truncation toward zero, like torch .long()   (combined_loss.py:97)
  // explicit _rn intrinsics: no FMA contraction, so the rounding matches torch's separate mul and sub
  row = static_cast<int>(__fsub_rn(__fmul_rn(static_cast<float>(p.coords[2 * st]) + 0.5f, p.coord_scale), 0.5f));
  col = static_cast<int>(__fsub_rn(__fmul_rn(static_cast<float>(p.coords[2 * st + 1]) + 0.5f, p.coord_scale), 0.5f));
  return row >= 0 && row < p.Hs && col >= 0 && col < p.Ws;
}
__device__ __forceinline__ float sample_weight(int mode, float obs) {
  if (mode == 1) return 1.f + log1pf(obs);
  if (mode == 2) return 1.f + sqrtf(obs);
  if (mode == 3) return obs >= 50.f ? 5.f : (obs >= 25.f ? 3.f : (obs >= 10.f ? 2.f : 1.f));
  return 1.f;
}

// pass C1: weighted L1 at the stations + number of valid observations
__global__ void __launch_bounds__(256) loss_point_kernel(const LossParams p) {
  __shared__ float red[8];
  const size_t n = static_cast<size_t>(p.B) * p.T * p.n_st;
  float acc = 0.f, cnt = 0.f;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int st = i % p.n_st;
    const size_t bt = i / p.n_st;
    int row, col;
    if (!station_pixel(p, st, row, col)) continue;
    const float obs = p.svals_has_batch ? p.svals[i] : p.svals[(bt % p.T) * p.n_st + st];
    if (isnan(obs)) continue;
    const float pv = p.pred[(bt * p.Hs + row) * p.Ws + col];
    acc += fabsf(pv - obs) * sample_weight(p.weight_mode, obs);
    cnt += 1.f;
  }
  float t = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(p.sums + 4, t);
  t = block_sum(cnt, red);
  if (threadIdx.x == 0) atomicAdd(p.sums + 5, t);
}
// pass C2: scatter the station-term gradient (needs the valid count from C1)
__global__ void __launch_bounds__(256) loss_point_grad_kernel(const LossParams p) {
  const size_t n = static_cast<size_t>(p.B) * p.T * p.n_st;
  const float cnt = p.sums[5];
  if (cnt <= 0.f) return;
  const float c = (p.grad_scale ? *p.grad_scale : 1.f) * p.l_point / cnt;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int st = i % p.n_st;
    const size_t bt = i / p.n_st;
    int row, col;
    if (!station_pixel(p, st, row, col)) continue;
    const float obs = p.svals_has_batch ? p.svals[i] : p.svals[(bt % p.T) * p.n_st + st];
    if (isnan(obs)) continue;
    const size_t pi = (bt * p.Hs + row) * p.Ws + col;
    atomicAdd(p.dpred + pi, c * sample_weight(p.weight_mode, obs) * sgn(p.pred[pi] - obs));
  }
}

// terms_out[5] = total, point, conserve, smooth, temporal (device side, no host sync)
__global__ void loss_finalize_kernel(const LossParams p, float* terms) {
  const float n_bt = static_cast<float>(p.B) * p.T;
  const float hw = static_cast<float>(p.Hs) * p.Ws;
  const float point = p.sums[5] > 0.f ? p.sums[4] / p.sums[5] : 0.f;
  const float cons = p.sums[0] / (n_bt * p.H * p.W);
  // torch: the mean of an empty tensor is NaN (Ws == 1, Hs == 1 or T == 1), kept as the reference has it
  const float smooth = p.sums[1] / (n_bt * p.Hs * (p.Ws - 1)) + p.sums[2] / (n_bt * (p.Hs - 1) * p.Ws);
  const float temp = p.sums[3] / (static_cast<float>(p.B) * (p.T - 1) * hw);
  terms[0] = p.l_point * point + p.l_cons * cons + p.l_smooth * smooth + p.l_temp * temp;
  terms[1] = point;
  terms[2] = cons;
  terms[3] = smooth;
  terms[4] = temp;
}

}  // namespace plc
