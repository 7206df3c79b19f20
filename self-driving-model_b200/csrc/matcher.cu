// Hungarian matcher: batched cost-matrix kernel (device) + rectangular LSAP (host).
// Replaces the per-image Python loop of training/hungarian_matcher.py:34-85
// (softmax, gather, cdist, box_convert x2, generalized_box_iou, 3 axpy, .cpu()).
#include <math.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "common.cuh"

constexpr int MQ_TILE = 128;  // queries per CTA (the per-image target staging is amortised over 128 rows)

// Non-contracted arithmetic: the reference evaluates each step as a separate rounded
// fp32 elementwise op, so no FMA contraction here.
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }

// xyxy corners from a D=4 (cx,cy,w,h) box (torchvision box_convert) or a D=7
// (cx,cy,cz,w,l,h,yaw) box (BEV: x,y,w,l; hungarian_matcher.py:56-64)
__device__ __forceinline__ void box_corners(const float* b, int D, float& x1, float& y1, float& x2, float& y2) {
  float cx = b[0], cy = b[1];
  float w = (D == 4) ? b[2] : b[3];
  float h = (D == 4) ? b[3] : b[4];
  float hw = fmul(0.5f, w), hh = fmul(0.5f, h);
  x1 = fsub(cx, hw);
  y1 = fsub(cy, hh);
  x2 = fadd(cx, hw);
  y2 = fadd(cy, hh);
}

// One CTA = 128 queries of one image.  Everything a (query, target) pair needs is staged in shared memory first
// (class probabilities of the 32 queries; per-box corners and areas, computed ONCE per box instead of once per
// pair), then warp w sweeps rows w, w+8, ... with its lanes along the target axis: conflict-free shared-memory
// reads, no integer division, and every store instruction writes one contiguous run of the padded cost row.
// The arithmetic per pair is the same rounded fp32 op sequence as before (= torch's eager ops).
struct BoxPre { float x1, y1, x2, y2, area; };

__device__ __forceinline__ BoxPre box_pre(const float* b, int D) {
  BoxPre r;
  box_corners(b, D, r.x1, r.y1, r.x2, r.y2);
  r.area = fmul(fsub(r.x2, r.x1), fsub(r.y2, r.y1));
  return r;
}

__global__ __launch_bounds__(256) void hungarian_cost_kernel(
    const float* __restrict__ logits, const float* __restrict__ boxes,
    const float* __restrict__ tgt_boxes, const int64_t* __restrict__ tgt_labels,
    const int32_t* __restrict__ n_tgt, float* __restrict__ cost, int Q, int C, int D, int Nmax,
    float w_class, float w_bbox, float w_giou) {
  extern __shared__ float sm[];
  const int b = blockIdx.y;
  const int q0 = blockIdx.x * MQ_TILE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = min(n_tgt[b], Nmax);
  const bool giou_on = (w_giou > 0.f) && (D == 4 || D == 7);
  // shared-memory carve-up (floats): probabilities [32][C] | query boxes [32][D] | query pre [5][32] |
  //                                  target boxes [D][Nmax] | target pre [5][Nmax] | target label [Nmax] (int)
  float* sprob = sm;
  float* sqb = sprob + MQ_TILE * C;
  float* sqp = sqb + MQ_TILE * D;
  float* stb = sqp + 5 * MQ_TILE;
  float* stp = stb + D * Nmax;
  int* slab = reinterpret_cast<int*>(stp + 5 * Nmax);
  for (int n = threadIdx.x; n < nt; n += blockDim.x) {
    const float* tb = tgt_boxes + ((int64_t)b * Nmax + n) * D;
    for (int k = 0; k < D; ++k) stb[k * Nmax + n] = tb[k];
    if (giou_on) {
      const BoxPre r = box_pre(tb, D);
      stp[n] = r.x1; stp[Nmax + n] = r.y1; stp[2 * Nmax + n] = r.x2; stp[3 * Nmax + n] = r.y2; stp[4 * Nmax + n] = r.area;
    }
    int64_t lab = tgt_labels[(int64_t)b * Nmax + n];
    if (lab < 0) lab += C;  // torch negative indexing
    slab[n] = (lab >= 0 && lab < C) ? (int)lab : -1;
  }
  if (threadIdx.x < MQ_TILE && q0 + threadIdx.x < Q) {
    const int r = threadIdx.x;
    const float* pb = boxes + ((int64_t)b * Q + q0 + r) * D;
    for (int k = 0; k < D; ++k) sqb[r * D + k] = pb[k];
    if (giou_on) {
      const BoxPre v = box_pre(pb, D);
      sqp[r] = v.x1; sqp[MQ_TILE + r] = v.y1; sqp[2 * MQ_TILE + r] = v.x2; sqp[3 * MQ_TILE + r] = v.y2; sqp[4 * MQ_TILE + r] = v.area;
    }
  }
  // softmax over classes, one warp per query row
  for (int r = warp; r < MQ_TILE; r += (blockDim.x >> 5)) {
    int q = q0 + r;
    if (q >= Q) continue;
    const float* lr = logits + ((int64_t)b * Q + q) * C;
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, lr[c]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int c = lane; c < C; c += 32) {
      float e = expf(lr[c] - mx);
      sprob[r * C + c] = e;
      s += e;
    }
    s = warp_sum(s);
    for (int c = lane; c < C; c += 32) sprob[r * C + c] = sprob[r * C + c] / s;
  }
  __syncthreads();
  for (int r = warp; r < MQ_TILE; r += (blockDim.x >> 5)) {
    const int q = q0 + r;
    if (q >= Q) break;
    float* crow = cost + ((int64_t)b * Q + q) * Nmax;
    float ax1 = 0.f, ay1 = 0.f, ax2 = 0.f, ay2 = 0.f, area1 = 0.f;
    if (giou_on) { ax1 = sqp[r]; ay1 = sqp[MQ_TILE + r]; ax2 = sqp[2 * MQ_TILE + r]; ay2 = sqp[3 * MQ_TILE + r]; area1 = sqp[4 * MQ_TILE + r]; }
    for (int n = lane; n < Nmax; n += 32) {
      float out = 0.f;
      if (n < nt) {
        const int lab = slab[n];
        const float p = lab >= 0 ? sprob[r * C + lab] : __int_as_float(0x7fc00000);
        float l1 = 0.f;
        for (int k = 0; k < D; ++k) l1 = fadd(l1, fabsf(fsub(sqb[r * D + k], stb[k * Nmax + n])));
        float g = 0.f;
        if (giou_on) {
          const float bx1 = stp[n], by1 = stp[Nmax + n], bx2 = stp[2 * Nmax + n], by2 = stp[3 * Nmax + n], area2 = stp[4 * Nmax + n];
          float iw = fmaxf(fsub(fminf(ax2, bx2), fmaxf(ax1, bx1)), 0.f);
          float ih = fmaxf(fsub(fminf(ay2, by2), fmaxf(ay1, by1)), 0.f);
          float inter = fmul(iw, ih);
          float uni = fsub(fadd(area1, area2), inter);
          float iou = __fdiv_rn(inter, uni);
          float cw = fmaxf(fsub(fmaxf(ax2, bx2), fminf(ax1, bx1)), 0.f);
          float ch = fmaxf(fsub(fmaxf(ay2, by2), fminf(ay1, by1)), 0.f);
          float areai = fmul(cw, ch);
          g = fsub(iou, __fdiv_rn(fsub(areai, uni), areai));
        }
        // C = w_bbox*cost_bbox + w_class*(-prob) + w_giou*(-giou)   (hungarian_matcher.py:73-75)
        out = fadd(fadd(fmul(w_bbox, l1), fmul(w_class, -p)), fmul(w_giou, -g));
      }
      crow[n] = out;
    }
  }
}

extern "C" int amoe_hungarian_cost_fwd(amoe_ctx* ctx, const float* logits, const float* boxes,
                                       const float* tgt_boxes, const int64_t* tgt_labels,
                                       const int32_t* n_tgt, float* cost, int B, int Q, int C, int D,
                                       int Nmax, float w_class, float w_bbox, float w_giou,
                                       void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && logits && boxes && n_tgt && cost, "amoe_hungarian_cost_fwd: NULL argument");
  AMOE_REQUIRE(Nmax == 0 || (tgt_boxes && tgt_labels), "amoe_hungarian_cost_fwd: NULL targets");
  AMOE_REQUIRE(D >= 1 && D <= 16, "amoe_hungarian_cost_fwd: box dimension %d out of range", D);
  if (B == 0 || Q == 0 || Nmax == 0) return 0;
  size_t smem = ((size_t)MQ_TILE * (C + D + 5) + (size_t)Nmax * (D + 6)) * sizeof(float);
  AMOE_REQUIRE(smem <= 200 * 1024, "amoe_hungarian_cost_fwd: %d classes x %d targets do not fit in shared memory", C, Nmax);
  if (smem > 48 * 1024)
    AMOE_CHECK_CUDA(cudaFuncSetAttribute(hungarian_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(ceil_div(Q, MQ_TILE), B);
  hungarian_cost_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(
      logits, boxes, tgt_boxes, tgt_labels, n_tgt, cost, Q, C, D, Nmax, w_class, w_bbox, w_giou);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

// ---------------------------------------------------------------------------------
// Host LSAP: shortest-augmenting-path algorithm for the rectangular assignment
// problem (D. F. Crouse, "On implementing 2D rectangular assignment algorithms",
// IEEE T-AES 2016) — the algorithm behind scipy.optimize.linear_sum_assignment
// (scipy 1.18.1, the reference's call at hungarian_matcher.py:79).  Same scan order
// (remaining columns kept in reverse order, swap-remove), same tie rule (prefer an
// unassigned column among equal minima), double arithmetic, so assignments agree.
// ---------------------------------------------------------------------------------
namespace {

struct LsapWork {
  std::vector<double> cost, u, v, sp;
  std::vector<int> path, col4row, row4col, remaining;
  std::vector<char> SR, SC;
};

// cost: nr x nc row-major with nr <= nc.  Returns 0 ok, -2 invalid entry, -3 infeasible.
int lsap_solve(int nr, int nc, LsapWork& wk) {
  const double* cost = wk.cost.data();
  for (int64_t i = 0; i < (int64_t)nr * nc; ++i)
    if (cost[i] != cost[i] || cost[i] == -INFINITY) return -2;
  wk.u.assign(nr, 0.0);
  wk.v.assign(nc, 0.0);
  wk.sp.resize(nc);
  wk.path.assign(nc, -1);
  wk.col4row.assign(nr, -1);
  wk.row4col.assign(nc, -1);
  wk.remaining.resize(nc);
  wk.SR.resize(nr);
  wk.SC.resize(nc);
  for (int cur = 0; cur < nr; ++cur) {
    // --- shortest augmenting path from row `cur` ---
    double min_val = 0.0;
    int n_rem = nc;
    for (int t = 0; t < nc; ++t) wk.remaining[t] = nc - t - 1;
    std::fill(wk.SR.begin(), wk.SR.end(), 0);
    std::fill(wk.SC.begin(), wk.SC.end(), 0);
    std::fill(wk.sp.begin(), wk.sp.end(), INFINITY);
    int sink = -1, i = cur;
    while (sink == -1) {
      int index = -1;
      double lowest = INFINITY;
      wk.SR[i] = 1;
      const double* crow = cost + (int64_t)i * nc;
      const double ui = wk.u[i];
      for (int t = 0; t < n_rem; ++t) {
        int j = wk.remaining[t];
        double r = min_val + crow[j] - ui - wk.v[j];
        if (r < wk.sp[j]) {
          wk.path[j] = i;
          wk.sp[j] = r;
        }
        if (wk.sp[j] < lowest || (wk.sp[j] == lowest && wk.row4col[j] == -1)) {
          lowest = wk.sp[j];
          index = t;
        }
      }
      min_val = lowest;
      if (min_val == INFINITY) return -3;
      int j = wk.remaining[index];
      if (wk.row4col[j] == -1) sink = j;
      else i = wk.row4col[j];
      wk.SC[j] = 1;
      wk.remaining[index] = wk.remaining[--n_rem];
    }
    // --- dual update ---
    wk.u[cur] += min_val;
    for (int r = 0; r < nr; ++r)
      if (wk.SR[r] && r != cur) wk.u[r] += min_val - wk.sp[wk.col4row[r]];
    for (int j = 0; j < nc; ++j)
      if (wk.SC[j]) wk.v[j] -= min_val - wk.sp[j];
    // --- augment ---
    int j = sink;
    while (true) {
      int r = wk.path[j];
      wk.row4col[j] = r;
      std::swap(wk.col4row[r], j);
      if (r == cur) break;
    }
  }
  return 0;
}

int lsap_one(const float* cost, int Q, int Nmax, int nt, int64_t* rows, int64_t* cols, LsapWork& wk) {
  // problem is Q x nt; the solver wants nr <= nc, so transpose when Q > nt
  if (nt == 0 || Q == 0) return 0;
  const bool transpose = nt < Q;
  const int nr = transpose ? nt : Q, nc = transpose ? Q : nt;
  wk.cost.resize((size_t)nr * nc);
  if (transpose) {
    for (int q = 0; q < Q; ++q)
      for (int n = 0; n < nt; ++n) wk.cost[(size_t)n * Q + q] = (double)cost[(size_t)q * Nmax + n];
  } else {
    for (int q = 0; q < Q; ++q)
      for (int n = 0; n < nt; ++n) wk.cost[(size_t)q * nt + n] = (double)cost[(size_t)q * Nmax + n];
  }
  int rc = lsap_solve(nr, nc, wk);
  if (rc != 0) return rc;
  if (transpose) {
    // pairs (query = col4row[t], target = t) ordered by query index
    std::vector<int> order(nr);
    for (int t = 0; t < nr; ++t) order[t] = t;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return wk.col4row[a] < wk.col4row[b]; });
    for (int k = 0; k < nr; ++k) {
      rows[k] = wk.col4row[order[k]];
      cols[k] = order[k];
    }
  } else {
    for (int q = 0; q < nr; ++q) {
      rows[q] = q;
      cols[q] = wk.col4row[q];
    }
  }
  return 0;
}

}  // namespace

extern "C" int amoe_lsap_batched_host(const float* cost_host, const int32_t* n_tgt_host, int B,
                                      int Q, int Nmax, int64_t* rows_host, int64_t* cols_host,
                                      int32_t* n_match_host, int n_threads) {
  AMOE_REQUIRE(n_tgt_host && rows_host && cols_host && n_match_host, "amoe_lsap_batched_host: NULL argument");
  AMOE_REQUIRE(cost_host || B == 0 || Q == 0 || Nmax == 0, "amoe_lsap_batched_host: NULL cost");
  const int K = std::min(Q, Nmax);
  for (int b = 0; b < B; ++b) {
    AMOE_REQUIRE(n_tgt_host[b] >= 0 && n_tgt_host[b] <= Nmax, "amoe_lsap_batched_host: n_tgt[%d]=%d out of [0,%d]", b, n_tgt_host[b], Nmax);
    n_match_host[b] = std::min(Q, (int)n_tgt_host[b]);
  }
  if (B == 0) return 0;
  n_threads = std::max(1, std::min(n_threads, B));
  std::vector<int> status(B, 0);
  auto work = [&](int tid) {
    LsapWork wk;
    for (int b = tid; b < B; b += n_threads)
      status[b] = lsap_one(cost_host + (size_t)b * Q * Nmax, Q, Nmax, n_tgt_host[b],
                           rows_host + (size_t)b * K, cols_host + (size_t)b * K, wk);
  };
  if (n_threads == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
    for (auto& t : th) t.join();
  }
  for (int b = 0; b < B; ++b) {
    if (status[b] == -2) {
      amoe_set_error("matrix contains invalid numeric entries (image %d)", b);
      return -2;
    }
    if (status[b] == -3) {
      amoe_set_error("cost matrix is infeasible (image %d)", b);
      return -3;
    }
  }
  return 0;
}
