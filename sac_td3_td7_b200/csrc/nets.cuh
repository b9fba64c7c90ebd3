// nets.cuh -- packed parameter layouts of the reference networks and small builders that turn
// "layer L applied to these activations" into GemmProblem entries.
//
// Layouts (rl/nn/sale.py:32-39,72-75,101-104; rl/nn/mlp.py:24-35): every nn.Linear is stored as
// W[out][ld] (ld = in rounded up to 4 floats, padding zero) followed by b[out]; layer offsets are
// multiples of 4 floats so float4 access is always aligned.  A critic's first layer reads cat[state, action]:
// its action columns start at the next multiple of 4 after the state columns (`split`, zero padding in between), so
// that both column blocks are 16-byte aligned operands (TMA, 128-bit loads) whatever the observation width.
// A network is ONE contiguous blob, so hard target copies are a single device memcpy and Polyak / Adam run over
// one range.
#pragma once
#include "host.cuh"

namespace b200rl {

enum NetKind : int { kNetSaleEncoder = 0, kNetSaleActor = 1, kNetSaleCritic = 2, kNetMlp = 3 };

struct LayerDesc {
  long long w_off, b_off;
  int out, in, ld;
  int split;  // logical input column where a second, 4-float aligned column block starts (0: one dense block)
  // physical column of logical input column c
  int phys(int c) const { return (split && c >= split) ? c + ((int)round_up(split, 4) - split) : c; }
};

struct NetLayout {
  std::vector<LayerDesc> layers;
  long long numel = 0;
  void add(int in, int out, int split = 0) {
    LayerDesc d;
    if (split <= 0 || split >= in || (split & 3) == 0) split = 0;
    d.in = in; d.out = out; d.split = split;
    d.ld = split ? (int)(round_up(split, 4) + round_up(in - split, 4)) : (int)round_up(in, 4);
    d.w_off = numel;
    numel += (long long)out * d.ld;
    d.b_off = numel;
    numel += round_up(out, 4);
    layers.push_back(d);
  }
};

// mlp_in / mlp_out only for kNetMlp
inline NetLayout make_layout(int kind, int S, int A, int H, int Z, int mlp_in = 0, int mlp_out = 0, int mlp_split = 0) {
  NetLayout L;
  switch (kind) {
    case kNetSaleEncoder:  // zs1 zs2 zs3 zsa1 zsa2 zsa3
      L.add(S, H); L.add(H, H); L.add(H, Z); L.add(Z + A, H); L.add(H, H); L.add(H, Z);
      break;
    case kNetSaleActor:  // l0 l1 l2 l3
      L.add(S, H); L.add(Z + H, H); L.add(H, H); L.add(H, A);
      break;
    case kNetSaleCritic:  // q01 q1 q2 q3
      L.add(S + A, H, S); L.add(2 * Z + H, H); L.add(H, H); L.add(H, 1);
      break;
    default:  // mlp.0 mlp.2 mlp.4
      L.add(mlp_in, H, mlp_split); L.add(H, H); L.add(H, mlp_out);
      break;
  }
  return L;
}

// activation view: [rows][cols] with leading dimension ld; gs = stride between stacked seeds
struct Mat {
  float* p = nullptr;
  int rows = 0, cols = 0, ld = 0;
  long long gs = 0;
  Mat rows_from(int r0, int n) const { Mat m = *this; m.p = p + (long long)r0 * ld; m.rows = n; return m; }
};

// a network instance: parameter blob (+ Adam state and optional gradient blob when trained)
struct Net {
  NetLayout L;
  float* p = nullptr;
  float* m = nullptr;
  float* v = nullptr;
  float* grad = nullptr;
  long long gs = 0;              // blob stride between stacked seeds (= L.numel)
  const AdamScalars* adam = nullptr;
};

// ---- problem builders -------------------------------------------------------------------------
inline GemmProblem fwd_problem(const Net& n, int layer, std::initializer_list<Mat> xs, const Mat& y, int act) {
  const LayerDesc& d = n.L.layers[layer];
  GemmProblem pr = make_problem(y.rows, d.out, y.p, y.ld, y.gs);
  int koff = 0;
  for (const Mat& x : xs) {
    add_seg(pr, x.p, x.ld, x.gs, n.p + d.w_off + d.phys(koff), d.ld, n.gs, x.cols);  // (a segment never straddles `split`)
    koff += x.cols;
  }
  pr.bias = n.p + d.b_off;
  pr.bias_gs = n.gs;
  pr.act = act;
  return pr;
}

// dX[:, 0:ncols] (+)= dZ . W[:, k_lo : k_lo+ncols]; optionally multiplied by act'(H)
struct DgradTerm { const Net* net; int layer; Mat dZ; int k_lo; };
inline GemmProblem dgrad_problem(std::initializer_list<DgradTerm> terms, int ncols, const Mat& dX, const Mat* H, int act) {
  GemmProblem pr = make_problem(dX.rows, ncols, dX.p, dX.ld, dX.gs);
  for (const DgradTerm& t : terms) {
    const LayerDesc& d = t.net->L.layers[t.layer];
    add_seg(pr, t.dZ.p, t.dZ.ld, t.dZ.gs, t.net->p + d.w_off + d.phys(t.k_lo), d.ld, t.net->gs, d.out);
  }
  if (H) { pr.aux = H->p; pr.aux_ld = H->ld; pr.aux_gs = H->gs; pr.act = act; }
  return pr;
}

// dW[:, cols of each input segment] = dZ^T . X_seg ; db = column sums of dZ ; Adam fused
inline void wgrad_problems(std::vector<GemmProblem>& out, const Net& n, int layer, const Mat& dZ, std::initializer_list<Mat> xs, bool with_adam) {
  const LayerDesc& d = n.L.layers[layer];
  int koff = 0;
  auto finish = [&](GemmProblem& pr, long long off, int ld) {
    if (n.grad) { pr.C = n.grad + off; pr.c_ld = ld; pr.c_gs = n.gs; }
    if (with_adam) {
      pr.P = n.p + off; pr.M = n.m + off; pr.V = n.v + off;
      pr.p_ld = ld; pr.p_gs = n.gs; pr.adam = n.adam;
    }
  };
  for (const Mat& x : xs) {
    GemmProblem pr = make_problem(d.out, x.cols, nullptr, d.ld, 0);
    add_seg(pr, dZ.p, dZ.ld, dZ.gs, x.p, x.ld, x.gs, dZ.rows);
    finish(pr, d.w_off + d.phys(koff), d.ld);
    out.push_back(pr);
    koff += x.cols;
  }
  GemmProblem pb = make_problem(d.out, 1, nullptr, 1, 0);
  add_seg(pb, dZ.p, dZ.ld, dZ.gs, nullptr, 0, 0, dZ.rows);
  pb.seg[0].b_ones = 1;
  finish(pb, d.b_off, 1);
  out.push_back(pb);
}

}  // namespace b200rl
