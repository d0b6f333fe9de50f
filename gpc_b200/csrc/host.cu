// host.cu -- the native callers either side of the hot path: the scaled-conjugate-gradient loop
// that asks for evaluations (COptimisable::scgOptimise, COptimisable.cpp:246-396) and the SVM-light reader that feeds
// X and y (CClctrl::readSvmlDataFile, CClctrl.cpp:55-171).  Host code only; the evaluations go through gpc_eval.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "common.cuh"

using namespace gpc;

namespace {

const double HALFLOGTWOPI = 0.9189385332046727;  // ndlutil.h

// ---- the objective: -log p(y | X, theta) of an FTC GP in TRANSFORMED parameters ---------------------------------
struct GpObjective {
  gpc_ctx* ctx;
  gpc_kcomp* comps;
  int ncomp;
  int P;
  std::vector<int> tr;          // transform of every parameter (component order)
  std::vector<double*> slot;    // where the natural value of parameter i lives (caller's gpc_kcomp::params)
  std::vector<double> w_cached, g_cached;
  double obj_cached;
  bool have;
  int64_t N;
  int d;
  int evals;

  void set(const std::vector<double>& w) {
    for (int i = 0; i < P; i++) *slot[i] = gpc_transform_atox(tr[i], w[i]);
  }
  // one device evaluation gives the objective AND its gradient; asking for either at the same w again is free
  // (the reference factorises K twice per accepted SCG step: COptimisable.cpp:333 then :347)
  int eval(const std::vector<double>& w) {
    if (have && w == w_cached) return GPC_OK;
    set(w);
    double out[3];
    std::vector<double> gnat((size_t)P);
    int rc = gpc_eval(ctx, comps, ncomp, 0, out, gnat.data(), nullptr);
    if (rc != GPC_OK) return rc;
    evals++;
    const double ll = -0.5 * (out[1] + (double)d * out[0]) - (double)d * (double)N * HALFLOGTWOPI;  // CGp.cpp:1002-1013
    obj_cached = -ll;
    g_cached.resize((size_t)P);
    for (int i = 0; i < P; i++) g_cached[i] = -gnat[i] * gpc_transform_gradfact(tr[i], *slot[i]);  // CKern.cpp:50-63
    w_cached = w;
    have = true;
    return GPC_OK;
  }
};

// the optimiser sees any model through this pair (COptimisable::computeObjectiveGradParams, COptimisable.h:15-239)
struct CallbackObjective {
  gpc_objective_fn fn;
  void* user;
  int P;
  std::vector<double> w_cached, g_cached;
  double obj_cached;
  bool have;
  int evals;
  void set(const std::vector<double>&) {}
  int eval(const std::vector<double>& w) {
    if (have && w == w_cached) return GPC_OK;
    g_cached.resize((size_t)P);
    int rc = fn(user, w.data(), P, &obj_cached, g_cached.data());
    if (rc != GPC_OK) return rc;
    evals++;
    w_cached = w;
    have = true;
    return GPC_OK;
  }
};

double dot(const std::vector<double>& a, const std::vector<double>& b) {
  double s = 0.0;
  for (size_t i = 0; i < a.size(); i++) s += a[i] * b[i];
  return s;
}

// ---- SVM-light reader --------------------------------------------------------------------------------------------
// Same semantics as the reference's two-pass reader: one data point per line; tokens are separated by single spaces
// (a tab is part of a token); the first token is the label (atof), every further token is index:value with a 1-based
// index (atoi / atof, so trailing garbage is ignored as the C library does); a line starting with '#' is skipped; a
// trailing '\r' is dropped; an empty line is a data point with label 0 and no features; D = the largest index seen.
static int svml_load(const char* path, std::vector<char>& buf) {
  FILE* fp = fopen(path, "rb");
  if (!fp) {
    set_error(std::string("gpc_svml: cannot open ") + path);
    return GPC_ERR_ARG;
  }
  fseek(fp, 0, SEEK_END);
  long sz = ftell(fp);
  fseek(fp, 0, SEEK_SET);
  buf.resize((size_t)sz + 1);
  size_t got = sz > 0 ? fread(buf.data(), 1, (size_t)sz, fp) : 0;
  fclose(fp);
  buf.resize(got + 1);
  buf[got] = '\0';
  return GPC_OK;
}

// walks the buffer line by line; cb(line_begin, line_end) for every data line
template <class F>
static void svml_lines(std::vector<char>& buf, F cb) {
  char* p = buf.data();
  char* end = buf.data() + buf.size() - 1;
  while (p < end) {
    char* e = (char*)memchr(p, '\n', (size_t)(end - p));
    char* le = e ? e : end;
    char* stop = le;
    if (stop > p && stop[-1] == '\r') stop--;
    if (!(stop > p && p[0] == '#')) cb(p, stop);
    p = e ? e + 1 : end;
  }
}

}  // namespace

extern "C" {

int gpc_ctx_dims(gpc_ctx* c, int64_t* N, int* D, int* d);  // api.cu

// Scaled conjugate gradients (Moller 1993), step for step COptimisable::scgOptimise (COptimisable.cpp:246-396) including
// the two quirks that shape its trajectory: step 3 adds lambdaDiff*|p| (not |p|^2) to delta (:313) and the
// convergence test looks at CMatrix::max(), which as implemented is max(p[0], p[last]) (CMatrix.cpp:568-577).
// F: eval(w) -> rc, then obj_cached / g_cached hold objective and gradient at w (cached per point); set(w).
}  // extern "C"

namespace {
template <class F>
int scg_loop(F& f, std::vector<double>& w, int max_iters, double param_tol, double obj_tol, double* trace, int* iters_out) {
  const int nP = (int)w.size();
  std::vector<double> wPlus((size_t)nP), r((size_t)nP), p((size_t)nP), s((size_t)nP, 0.0), rp((size_t)nP);
  const double m_step = 1.0e-4, m_reg = 1.0;
  double lam = m_reg, lamBar = 0.0, delta = 0.0;
  bool success = true;
  int rc = f.eval(w);
  if (rc != GPC_OK) return rc;
  double oldObj = f.obj_cached, newObj = oldObj;
  for (int i = 0; i < nP; i++) p[i] = r[i] = -f.g_cached[i];
  int it = 0;
  for (it = 1; it <= max_iters; it++) {
    const double normp = sqrt(dot(p, p)), normp2 = normp * normp;
    if (success) {  // 2: second-order information along p by a finite difference of the gradient
      const double sigma = m_step / normp;
      for (int i = 0; i < nP; i++) wPlus[i] = w[i] + sigma * p[i];
      if ((rc = f.eval(wPlus)) != GPC_OK) break;
      for (int i = 0; i < nP; i++) s[i] = (f.g_cached[i] + r[i]) / sigma;
      delta = dot(s, p);
    }
    const double lamDiff = lam - lamBar;  // 3
    for (int i = 0; i < nP; i++) s[i] += lamDiff * p[i];
    delta += lamDiff * normp;
    if (delta <= 0.0) {  // 4
      const double dn = delta / normp2;
      for (int i = 0; i < nP; i++) s[i] += (lam - 2.0 * dn) * p[i];
      lamBar = 2.0 * (lam - dn);
      delta = lam * normp2 - delta;
      lam = lamBar;
    }
    const double mu = dot(p, r);  // 5
    const double alpha = mu / delta;
    for (int i = 0; i < nP; i++) wPlus[i] = w[i] + alpha * p[i];  // 6
    if ((rc = f.eval(wPlus)) != GPC_OK) break;
    newObj = f.obj_cached;
    const double Delta = 2.0 * delta * (oldObj - newObj) / (mu * mu);
    if (Delta >= 0.0) {  // 7: accept; the gradient at wPlus came with the objective
      w = wPlus;
      oldObj = newObj;
      for (int i = 0; i < nP; i++) rp[i] = -f.g_cached[i];
      lamBar = 0.0;
      success = true;
      if (it % nP == 0) {
        p = rp;
      } else {
        const double beta = (dot(rp, rp) - dot(r, rp)) / mu;
        for (int i = 0; i < nP; i++) p[i] = beta * p[i] + rp[i];
      }
      r = rp;
      if (Delta >= 0.75) lam *= 0.5;
      if (lam < 1e-15) lam = 1e-15;
    } else {
      lamBar = lam;
      success = false;
    }
    if (Delta < 0.25) lam *= 4.0;  // 8
    if (trace) trace[it - 1] = oldObj;
    const double pmax = p[0] > p[nP - 1] ? p[0] : p[nP - 1];  // CMatrix::max() as implemented
    if (success && fabs(pmax * alpha) < param_tol && fabs(newObj - oldObj) < obj_tol) {  // 9
      it++;
      break;
    }
  }
  f.set(w);  // leave the accepted point with the model
  if (iters_out) *iters_out = it - 1 > max_iters ? max_iters : it - 1;
  return rc;
}
}  // namespace

extern "C" {

// any objective through a callback (the GP-LVM's [kernel][X] parameters, tests): w holds the start point and receives
// the result
int gpc_scg_minimise(gpc_objective_fn fn, void* user, double* w, int n, int max_iters, double param_tol, double obj_tol,
                     double* trace, int* iters_out, int* evals_out) {
  if (!fn || !w || n < 1 || max_iters < 0) {
    set_error("gpc_scg_minimise: bad arguments");
    return GPC_ERR_ARG;
  }
  CallbackObjective f;
  f.fn = fn;
  f.user = user;
  f.P = n;
  f.have = false;
  f.evals = 0;
  f.obj_cached = 0.0;
  std::vector<double> wv(w, w + n);
  int rc = scg_loop(f, wv, max_iters, param_tol, obj_tol, trace, iters_out);
  for (int i = 0; i < n; i++) w[i] = wv[i];
  if (evals_out) *evals_out = f.evals;
  return rc;
}

// the kernel hyper-parameters of an FTC GP: CGp::optimise (CGp.cpp:1537-1553) with the default optimiser
int gpc_gp_optimise_scg(gpc_ctx* ctx, gpc_kcomp* comps, int ncomp, int max_iters, double param_tol, double obj_tol,
                        double* trace, int* iters_out, int* evals_out) {
  if (!ctx || !comps || ncomp < 1 || max_iters < 0) {
    set_error("gpc_gp_optimise_scg: bad arguments");
    return GPC_ERR_ARG;
  }
  GpObjective f;
  f.ctx = ctx;
  f.comps = comps;
  f.ncomp = ncomp;
  f.have = false;
  f.evals = 0;
  f.obj_cached = 0.0;
  int D = 0;
  if (gpc_ctx_dims(ctx, &f.N, &D, &f.d) != GPC_OK) return GPC_ERR_STATE;
  for (int c = 0; c < ncomp; c++) {
    if (comps[c].nparams != gpc_kern_nparams(comps[c].type, D) || !comps[c].params) {
      set_error("gpc_gp_optimise_scg: component parameter count does not match its type");
      return GPC_ERR_ARG;
    }
    for (int i = 0; i < comps[c].nparams; i++) {
      f.tr.push_back(gpc_kern_transform(comps[c].type, i));
      f.slot.push_back(const_cast<double*>(comps[c].params) + i);
    }
  }
  const int nP = f.P = (int)f.slot.size();
  std::vector<double> w((size_t)nP);
  for (int i = 0; i < nP; i++) w[i] = gpc_transform_xtoa(f.tr[i], *f.slot[i]);
  int rc = scg_loop(f, w, max_iters, param_tol, obj_tol, trace, iters_out);
  if (evals_out) *evals_out = f.evals;
  return rc;
}

int gpc_svml_dims(const char* path, int64_t* nrows, int* ncols) {
  if (!path || !nrows || !ncols) return GPC_ERR_ARG;
  std::vector<char> buf;
  int rc = svml_load(path, buf);
  if (rc != GPC_OK) return rc;
  int64_t n = 0;
  int maxFeat = 0;
  bool bad = false;
  svml_lines(buf, [&](char* b, char* e) {
    n++;
    bool labelRead = false;
    char* p = b;
    while (p < e) {
      char* t = p;
      while (p < e && *p != ' ') p++;
      if (p > t) {
        if (labelRead) {
          char* colon = (char*)memchr(t, ':', (size_t)(p - t));
          if (!colon) {
            bad = true;
          } else {
            std::string idx(t, colon);
            int f = atoi(idx.c_str());
            if (f > maxFeat) maxFeat = f;
          }
        } else {
          labelRead = true;
        }
      }
      p++;
    }
  });
  if (bad) {
    set_error(std::string("gpc_svml: feature token without ':' in ") + path);  // FileFormatError, CClctrl.cpp:88-92
    return GPC_ERR_ARG;
  }
  *nrows = n;
  *ncols = maxFeat;
  return GPC_OK;
}

int gpc_svml_read(const char* path, double* X, int64_t ldx, double* y, int64_t nrows, int ncols) {
  if (!path || !X || !y || nrows < 0 || ncols < 0 || ldx < nrows) return GPC_ERR_ARG;
  std::vector<char> buf;
  int rc = svml_load(path, buf);
  if (rc != GPC_OK) return rc;
  for (int j = 0; j < ncols; j++) memset(X + (size_t)j * ldx, 0, sizeof(double) * (size_t)nrows);
  memset(y, 0, sizeof(double) * (size_t)nrows);
  int64_t row = 0;
  bool bad = false;
  svml_lines(buf, [&](char* b, char* e) {
    if (row >= nrows) {
      bad = true;
      row++;
      return;
    }
    bool labelRead = false;
    char* p = b;
    while (p < e) {
      char* t = p;
      while (p < e && *p != ' ') p++;
      if (p > t) {
        std::string tok(t, p);
        if (labelRead) {
          size_t colon = tok.find(':');
          std::string idx = tok.substr(0, colon);  // npos -> the whole token, as substr(0, -1) does in the reference
          std::string val = colon == std::string::npos ? std::string() : tok.substr(colon + 1);
          int f = atoi(idx.c_str());
          if (f < 1 || f > ncols) {
            bad = true;
          } else {
            X[row + (size_t)(f - 1) * ldx] = atof(val.c_str());
          }
        } else {
          y[row] = atof(tok.c_str());
          labelRead = true;
        }
      }
      p++;
    }
    row++;
  });
  if (bad || row != nrows) {
    set_error(std::string("gpc_svml: file does not match the given dimensions or has an index out of range: ") + path);
    return GPC_ERR_ARG;
  }
  return GPC_OK;
}

}  // extern "C"
