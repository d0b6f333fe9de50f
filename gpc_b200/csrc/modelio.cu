// modelio.cu -- GP and GP-LVM model files ("on-disk formats"): the text formats `gp learn` /
// `gplvm learn` write and `gp display / gnuplot / relearn`, `gplvm display / gnuplot` read.  Host code only.  Follows,
// field by field,
//   CGplvm::writeParamsToStream / readParamsFromStream, writeGplvmToFile           CGplvm.cpp:761-921
//   CStreamInterface::toStream / fromStream, readStringFromStream, writeToStream   CNdlInterfaces.h:21-175
//   CGp::writeParamsToStream / readParamsFromStream                                CGp.cpp:1605-1666
//   CMatrix::writeParamsToStream / readParamsFromStream / toUnheadedStream          CMatrix.cpp:1057-1097, 1158-1172
//   CKern / CComponentKern / CPolyKern stream functions, readKernFromStream         CKern.cpp:15-26, 94-126, 2668-2705, 4192-4278
//   CNoise stream functions, readNoiseFromStream                                    CNoise.cpp:275-305, 1813-1836
//   ndlstrutil::getline (skips '#' lines), tokenise                                 ndlstrutil.cpp:7-36
// including the format's quirks, because files must round-trip with the reference's own tools:
//   * the outermost object's version is written with ios::fixed ("0.200000"); every nested object sets ios::scientific
//     on top of it, which libstdc++ treats as hexfloat: nested versions and all matrix entries are "%a" text;
//   * a matrix entry with integer value is written as that integer;
//   * on reading, an entry WITHOUT a '.' goes through atoi (CMatrix.cpp:1081-1085): "0x1p-2" (0.25) reads back as 0,
//     "inf" as 0.  The reader reproduces this (the parity tests pin it); gpc_gp_model_check_roundtrip tells a caller
//     which values of a model would be lost that way before it writes;
//   * priors: the reference's writer emits them in a form its own reader rejects (CDist.cpp:4-10 writes no
//     baseType/type, CDist.cpp:338-357 expects them), so a file with numPriors != 0 is an error here as it is there.
#include <errno.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "common.cuh"

using namespace gpc;

namespace {

struct Reader {
  FILE* f;
  std::string path, err;
  explicit Reader(const char* p) : f(fopen(p, "rb")), path(p) {}
  ~Reader() {
    if (f) fclose(f);
  }
  // std::getline + ndlstrutil::getline: next line not starting with '#'
  bool line(std::string& out) {
    for (;;) {
      out.clear();
      int c;
      bool any = false;
      while ((c = fgetc(f)) != EOF) {
        any = true;
        if (c == '\n') break;
        out.push_back((char)c);
      }
      if (!any) return false;
      if (!out.empty() && out[0] == '#') continue;
      return true;
    }
  }
  bool fail(const std::string& what) {
    if (err.empty()) err = "gpc_gp_model_read: " + what + " in " + path;
    return false;
  }
  // readStringFromStream (CNdlInterfaces.h:85-94): "name=value", exactly two '='-separated tokens
  bool field(const char* name, std::string& val) {
    std::string l;
    if (!line(l)) return fail(std::string("end of file when expecting field: ") + name);
    std::vector<std::string> tok;
    size_t last = l.find_first_not_of('=', 0), pos = l.find_first_of('=', last);
    while (pos != std::string::npos || last != std::string::npos) {
      tok.push_back(l.substr(last, pos - last));
      last = l.find_first_not_of('=', pos);
      pos = l.find_first_of('=', last);
    }
    if (tok.size() != 2 || tok[0] != name) return fail(std::string("error when expecting field: ") + name);
    val = tok[1];
    return true;
  }
  bool field_long(const char* name, long* v) {
    std::string s;
    if (!field(name, s)) return false;
    *v = atol(s.c_str());
    return true;
  }
  bool field_double(const char* name, double* v) {
    std::string s;
    if (!field(name, s)) return false;
    *v = atof(s.c_str());
    return true;
  }
  bool expect(const char* name, const char* want) {
    std::string s;
    if (!field(name, s)) return false;
    if (s != want) return fail(std::string("mismatch in field ") + name + ": found " + s + ", expected " + want);
    return true;
  }
  bool version() {  // readVersionFromStream (CNdlInterfaces.h:35-42), MINVERSION 0.2
    double v;
    if (!field_double("version", &v)) return false;
    if (v < 0.2) return fail("file version below 0.2");
    return true;
  }
  // CMatrix::fromStream: version + readParamsFromStream (CMatrix.cpp:1057-1088)
  bool matrix(std::vector<double>& vals, long* rows, long* cols) {
    if (!version() || !expect("baseType", "matrix") || !expect("type", "doubleMatrix")) return false;
    if (!field_long("numRows", rows) || !field_long("numCols", cols)) return false;
    if (*rows < 0 || *cols < 0 || (double)*rows * (double)*cols > 1e8) return fail("unreasonable matrix size");
    vals.assign((size_t)(*rows * *cols), 0.0);
    for (long i = 0; i < *rows; i++) {
      std::string l;
      if (!line(l)) return fail("incorrect number of rows in matrix");
      if (!l.empty() && l[l.size() - 1] == '\r') l.erase(l.size() - 1);
      std::vector<std::string> tok;
      size_t last = l.find_first_not_of(' ', 0), pos = l.find_first_of(' ', last);
      while (pos != std::string::npos || last != std::string::npos) {
        tok.push_back(l.substr(last, pos - last));
        last = l.find_first_not_of(' ', pos);
        pos = l.find_first_of(' ', last);
      }
      if ((long)tok.size() != *cols) return fail("incorrect number of columns in a matrix row");
      for (long j = 0; j < *cols; j++) {
        const std::string& t = tok[(size_t)j];
        // the reference's rule: a token without '.' is an integer (CMatrix.cpp:1081-1085)
        vals[(size_t)(i + j * *rows)] = (t.find('.') == std::string::npos) ? (double)atoi(t.c_str()) : atof(t.c_str());
      }
    }
    return true;
  }
};

int kern_type_of(const std::string& t) {
  if (t == "white") return GPC_KERN_WHITE;
  if (t == "bias") return GPC_KERN_BIAS;
  if (t == "rbf") return GPC_KERN_RBF;
  if (t == "rbfard") return GPC_KERN_RBFARD;
  if (t == "matern32") return GPC_KERN_MATERN32;
  if (t == "matern52") return GPC_KERN_MATERN52;
  if (t == "lin") return GPC_KERN_LIN;
  if (t == "poly") return GPC_KERN_POLY;
  return -1;
}
const char* kern_name_of(int type) {
  static const char* names[] = {"white", "bias", "rbf", "rbfard", "matern32", "matern52", "lin", "poly"};
  return (type >= 0 && type < 8) ? names[type] : nullptr;
}

// one non-compound kernel after its "type=" line: CKern::readParamsFromStream (CKern.cpp:4260-4278), CPolyKern's
// (CKern.cpp:2685-2705)
bool read_leaf(Reader& r, int type, gpc_kern_spec* m, int* poff) {
  if (m->ncomp >= GPC_MAX_COMPONENTS) return r.fail("too many kernel components");
  long inDim, nPar;
  if (!r.field_long("inputDim", &inDim) || !r.field_long("numParams", &nPar)) return false;
  double degree = 2.0;
  if (type == GPC_KERN_POLY && !r.field_double("degree", &degree)) return false;
  std::vector<double> par;
  long rows, cols;
  if (!r.matrix(par, &rows, &cols)) return false;
  if (nPar != gpc_kern_nparams(type, (int)inDim) || rows * cols != nPar)
    return r.fail("listed number of parameters does not match computed number of parameters");
  if (*poff + nPar > GPC_MAX_PARAMS) return r.fail("too many kernel parameters");
  long nPri;
  if (!r.field_long("numPriors", &nPri)) return false;
  if (nPri != 0) return r.fail("priors in a model file (the reference's reader rejects its writer's prior format)");
  const int c = m->ncomp++;
  m->type[c] = type;
  m->nparams[c] = (int)nPar;
  m->degree[c] = degree;
  for (long i = 0; i < nPar; i++) m->params[*poff + i] = par[(size_t)i];
  *poff += (int)nPar;
  if (m->input_dim == 0) m->input_dim = (int)inDim;
  return true;
}

// readKernFromStream (CKern.cpp:4192-4259)
bool read_kern(Reader& r, gpc_kern_spec* m) {
  std::string t;
  if (!r.version() || !r.expect("baseType", "kern") || !r.field("type", t)) return false;
  int poff = 0;
  m->ncomp = 0;
  m->input_dim = 0;
  if (t == "cmpnd") {  // CComponentKern::readParamsFromStream (CKern.cpp:94-113)
    long inDim, nPar, nKern;
    if (!r.field_long("inputDim", &inDim) || !r.field_long("numParams", &nPar) || !r.field_long("numKerns", &nKern)) return false;
    m->top_is_cmpnd = 1;
    for (long i = 0; i < nKern; i++) {
      std::string ct;
      if (!r.version() || !r.expect("baseType", "kern") || !r.field("type", ct)) return false;
      int type = kern_type_of(ct);
      if (type < 0) return r.fail("kernel type " + ct + " is outside the device path");
      if (!read_leaf(r, type, m, &poff)) return false;
    }
    m->input_dim = (int)inDim;
    return true;
  }
  int type = kern_type_of(t);
  if (type < 0) return r.fail("kernel type " + t + " is outside the device path");
  m->top_is_cmpnd = 0;
  return read_leaf(r, type, m, &poff);
}

// readNoiseFromStream + CNoise::readParamsFromStream (CNoise.cpp:1813-1836, 286-305)
bool read_noise(Reader& r, gpc_noise_spec* n) {
  std::string nt;
  long v, rows, cols;
  if (!r.version() || !r.expect("baseType", "noise") || !r.field("type", nt)) return false;
  if (nt != "probit" && nt != "ncnm" && nt != "gaussian" && nt != "ordered" && nt != "scale")
    return r.fail("unknown noise type " + nt);
  snprintf(n->type, sizeof n->type, "%s", nt.c_str());
  if (!r.field_long("outputDim", &v)) return false;
  n->output_dim = (int)v;
  if (!r.field_long("numParams", &v)) return false;
  if (v < 0 || v > (long)(sizeof n->params / sizeof n->params[0])) return r.fail("noise numParams out of range");
  n->nparams = (int)v;
  std::vector<double> vals;
  if (!r.matrix(vals, &rows, &cols)) return false;
  if (rows * cols != n->nparams) return r.fail("number of noise parameters in file does not match");
  for (int j = 0; j < n->nparams; j++) n->params[j] = vals[(size_t)j];
  return true;
}

// ---- writer -----------------------------------------------------------------------------------------------------
// `out << val` with ios::fixed | ios::scientific set = "%a"; CMatrix::toUnheadedStream prints integer values as int
// (CMatrix.cpp:1158-1172; (int)val of an out-of-range double is INT_MIN on x86, so those stay hexfloat)
void put_value(FILE* f, double v) {
  if (v >= -2147483648.0 && v < 2147483648.0 && (v - (double)(int)v) == 0.0)
    fprintf(f, "%d ", (int)v);
  else
    fprintf(f, "%a ", v);
}
const char* NESTED_VERSION = "version=0x1.999999999999ap-3\n";  // 0.2 through "%a"

void put_matrix(FILE* f, const double* v, long rows, long cols) {
  fputs(NESTED_VERSION, f);
  fprintf(f, "baseType=matrix\ntype=doubleMatrix\nnumRows=%ld\nnumCols=%ld\n", rows, cols);
  for (long i = 0; i < rows; i++) {
    for (long j = 0; j < cols; j++) put_value(f, v[i + j * rows]);
    fputc('\n', f);
  }
}

// the value the reference's reader returns for what put_value writes
double reread(double v) {
  char buf[64];
  if (v >= -2147483648.0 && v < 2147483648.0 && (v - (double)(int)v) == 0.0)
    snprintf(buf, sizeof buf, "%d", (int)v);
  else
    snprintf(buf, sizeof buf, "%a", v);
  return strchr(buf, '.') ? atof(buf) : (double)atoi(buf);
}

int check_kern_noise(const gpc_kern_spec* k, const gpc_noise_spec* n, const char* who) {
  if (k->ncomp < 1 || k->ncomp > GPC_MAX_COMPONENTS || (!k->top_is_cmpnd && k->ncomp != 1) || n->nparams < 0 ||
      n->nparams > (int)(sizeof n->params / sizeof n->params[0])) {
    set_error(std::string(who) + ": sizes out of range");
    return GPC_ERR_ARG;
  }
  int tot = 0;
  for (int c = 0; c < k->ncomp; c++) {
    if (!kern_name_of(k->type[c]) || k->nparams[c] != gpc_kern_nparams(k->type[c], k->input_dim)) {
      set_error(std::string(who) + ": kernel component type / parameter count");
      return GPC_ERR_ARG;
    }
    tot += k->nparams[c];
  }
  if (tot > GPC_MAX_PARAMS) {
    set_error(std::string(who) + ": too many kernel parameters");
    return GPC_ERR_ARG;
  }
  return GPC_OK;
}

int check_model(const gpc_gp_model* m, const char* who) {
  if (!m) {
    set_error(std::string(who) + ": null model");
    return GPC_ERR_ARG;
  }
  if (m->approx_type != 0) {
    set_error(std::string(who) + ": sparse approximations are not supported");
    return GPC_ERR_ARG;
  }
  if (m->output_dim < 1 || m->output_dim > GPC_MODEL_MAX_OUT) {
    set_error(std::string(who) + ": sizes out of range");
    return GPC_ERR_ARG;
  }
  return check_kern_noise(&m->kern, &m->noise, who);
}

// readKernFromStream's counterpart: CComponentKern / CKern / CPolyKern writeParamsToStream (CKern.cpp:114-126, 15-26,
// 2668-2684)
void put_kern(FILE* f, const gpc_kern_spec* k) {
  int tot = 0;
  for (int c = 0; c < k->ncomp; c++) tot += k->nparams[c];
  if (k->top_is_cmpnd) {
    fputs(NESTED_VERSION, f);
    fprintf(f, "baseType=kern\ntype=cmpnd\ninputDim=%d\nnumParams=%d\nnumKerns=%d\n", k->input_dim, tot, k->ncomp);
  }
  int poff = 0;
  for (int c = 0; c < k->ncomp; c++) {
    fputs(NESTED_VERSION, f);
    fprintf(f, "baseType=kern\ntype=%s\ninputDim=%d\nnumParams=%d\n", kern_name_of(k->type[c]), k->input_dim, k->nparams[c]);
    if (k->type[c] == GPC_KERN_POLY) {
      const double deg = k->degree[c];
      if ((deg - (double)(int)deg) == 0.0)
        fprintf(f, "degree=%d\n", (int)deg);
      else
        fprintf(f, "degree=%a\n", deg);
    }
    put_matrix(f, k->params + poff, 1, k->nparams[c]);
    fputs("numPriors=0\n", f);
    poff += k->nparams[c];
  }
}
// CNoise::writeParamsToStream (CNoise.cpp:275-285)
void put_noise(FILE* f, const gpc_noise_spec* n) {
  fputs(NESTED_VERSION, f);
  fprintf(f, "baseType=noise\ntype=%s\noutputDim=%d\nnumParams=%d\n", n->type, n->output_dim, n->nparams);
  put_matrix(f, n->params, 1, n->nparams);
}

}  // namespace

int gpc_gp_model_read(const char* path, gpc_gp_model* m) {
  if (!path || !m) {
    set_error("gpc_gp_model_read: null argument");
    return GPC_ERR_ARG;
  }
  memset(m, 0, sizeof *m);
  Reader r(path);
  if (!r.f) {
    set_error(std::string("gpc_gp_model_read: cannot open ") + path);
    return GPC_ERR_ARG;
  }
  bool ok = false;
  do {
    long v;
    // CStreamInterface::fromStream + CGp::readParamsFromStream (CGp.cpp:1605-1650)
    if (!r.version() || !r.expect("baseType", "dataModel") || !r.expect("type", "gp")) break;
    if (!r.field_long("numData", &v)) break;
    m->num_data = v;
    if (!r.field_long("outputDim", &v)) break;
    m->output_dim = (int)v;
    if (!r.field_long("inputDim", &v)) break;
    m->input_dim = (int)v;
    if (!r.field_long("sparseApproximation", &v)) break;
    m->approx_type = (int)v;
    if (!r.field_long("numActive", &v)) break;
    m->num_active = (unsigned int)v;
    if (m->approx_type != 0) {
      r.fail("sparse approximation models (DTC/FITC/PITC) are not supported");
      break;
    }
    if (m->output_dim < 1 || m->output_dim > GPC_MODEL_MAX_OUT) {
      r.fail("outputDim out of range");
      break;
    }
    if (!r.field_long("learnScale", &v)) break;
    m->learn_scale = v != 0;
    if (!r.field_long("learnBias", &v)) break;
    m->learn_bias = v != 0;
    std::vector<double> vals;
    long rows, cols;
    if (!r.matrix(vals, &rows, &cols)) break;
    if (rows * cols != m->output_dim) {
      r.fail("scale does not have outputDim entries");
      break;
    }
    for (int j = 0; j < m->output_dim; j++) m->scale[j] = vals[(size_t)j];
    if (!r.matrix(vals, &rows, &cols)) break;
    if (rows * cols != m->output_dim) {
      r.fail("bias does not have outputDim entries");
      break;
    }
    for (int j = 0; j < m->output_dim; j++) m->bias[j] = vals[(size_t)j];
    if (!read_kern(r, &m->kern)) break;
    if (!read_noise(r, &m->noise)) break;
    ok = true;
  } while (false);
  if (!ok) {
    set_error(r.err.empty() ? std::string("gpc_gp_model_read: failed on ") + path : r.err);
    return GPC_ERR_ARG;
  }
  return GPC_OK;
}

int gpc_gp_model_write(const char* path, const gpc_gp_model* m, const char* comment) {
  int rc = check_model(m, "gpc_gp_model_write");
  if (rc != GPC_OK) return rc;
  if (!path) {
    set_error("gpc_gp_model_write: null path");
    return GPC_ERR_ARG;
  }
  FILE* f = fopen(path, "wb");
  if (!f) {
    set_error(std::string("gpc_gp_model_write: cannot open ") + path);
    return GPC_ERR_ARG;
  }
  // CStreamInterface::toFile / toStream (CNdlInterfaces.h:24-33, 151-159), CGp::writeParamsToStream (CGp.cpp:1653-1682)
  if (comment && comment[0]) fprintf(f, "# %s\n", comment);
  fputs("version=0.200000\n", f);
  fprintf(f, "baseType=dataModel\ntype=gp\nnumData=%lld\noutputDim=%d\ninputDim=%d\n", (long long)m->num_data, m->output_dim,
          m->input_dim);
  fprintf(f, "sparseApproximation=%d\nnumActive=%u\nlearnScale=%d\nlearnBias=%d\n", m->approx_type, m->num_active,
          m->learn_scale ? 1 : 0, m->learn_bias ? 1 : 0);
  put_matrix(f, m->scale, 1, m->output_dim);
  put_matrix(f, m->bias, 1, m->output_dim);
  put_kern(f, &m->kern);
  put_noise(f, &m->noise);
  const bool bad = ferror(f) != 0;
  if (fclose(f) != 0 || bad) {
    set_error(std::string("gpc_gp_model_write: write error on ") + path);
    return GPC_ERR_ARG;
  }
  return GPC_OK;
}

int gpc_gp_model_check_roundtrip(const gpc_gp_model* m, int* nlost, double* first_lost) {
  int rc = check_model(m, "gpc_gp_model_check_roundtrip");
  if (rc != GPC_OK) return rc;
  int lost = 0;
  double first = 0.0;
  auto see = [&](double v) {
    const double back = reread(v);
    if (!(back == v) && !(v != v && back != back)) {
      if (!lost) first = v;
      lost++;
    }
  };
  for (int j = 0; j < m->output_dim; j++) see(m->scale[j]);
  for (int j = 0; j < m->output_dim; j++) see(m->bias[j]);
  int tot = 0;
  for (int c = 0; c < m->kern.ncomp; c++) tot += m->kern.nparams[c];
  for (int i = 0; i < tot; i++) see(m->kern.params[i]);
  for (int i = 0; i < m->noise.nparams; i++) see(m->noise.params[i]);
  if (nlost) *nlost = lost;
  if (first_lost) *first_lost = first;
  return GPC_OK;
}

// ---- GP-LVM model files --------------------------------------------------------------------------------------------
int gpc_gplvm_model_read(const char* path, gpc_gplvm_model* m, double* Y, int64_t ldy, double* X, int64_t ldx, int* labels) {
  if (!path || !m) {
    set_error("gpc_gplvm_model_read: null argument");
    return GPC_ERR_ARG;
  }
  memset(m, 0, sizeof *m);
  Reader r(path);
  if (!r.f) {
    set_error(std::string("gpc_gplvm_model_read: cannot open ") + path);
    return GPC_ERR_ARG;
  }
  bool ok = false;
  do {
    long v;
    // CStreamInterface::fromStream + CGplvm::readParamsFromStream (CGplvm.cpp:802-898)
    if (!r.version() || !r.expect("baseType", "dataModel") || !r.expect("type", "gplvm")) break;
    if (!r.field_long("numData", &v)) break;
    m->num_data = v;
    if (!r.field_long("outputDim", &v)) break;
    m->output_dim = (int)v;
    if (!r.field_long("inputDim", &v)) break;
    m->latent_dim = (int)v;
    if (!r.field_long("latentRegularised", &v)) break;
    m->latent_regularised = v != 0;
    if (!r.field_long("backConstrained", &v)) break;
    m->back_constrained = v != 0;
    if (!r.field_long("dynamicsLearnt", &v)) break;
    m->dynamics_learnt = v != 0;
    if (m->back_constrained || m->dynamics_learnt) {
      r.fail("GP-LVM models with back constraints or dynamics are not supported");
      break;
    }
    if (m->num_data < 0 || m->output_dim < 1 || m->output_dim > GPC_MODEL_MAX_OUT || m->latent_dim < 1) {
      r.fail("sizes out of range");
      break;
    }
    if (!read_kern(r, &m->kern)) break;
    if (!read_noise(r, &m->noise)) break;
    // "Y:d,X:q[,labels:1]" (CGplvm.cpp:843-872)
    std::string l;
    if (!r.line(l)) {
      r.fail("end of file when expecting the Y:,X: line");
      break;
    }
    bool bad = false;
    size_t pos = 0;
    while (pos <= l.size() && !bad) {
      size_t e = l.find(',', pos);
      if (e == std::string::npos) e = l.size();
      std::string tok = l.substr(pos, e - pos);
      pos = e + 1;
      if (tok.empty()) continue;
      size_t c = tok.find(':');
      std::string key = tok.substr(0, c);
      int val = (c == std::string::npos) ? 0 : atoi(tok.c_str() + c + 1);
      if (key == "Y")
        bad = val != m->output_dim;
      else if (key == "X")
        bad = val != m->latent_dim;
      else if (key == "labels") {
        m->has_labels = 1;
        bad = val != 1;
      } else
        bad = true;
    }
    if (bad) {
      r.fail("file format error in the Y:,X: line");
      break;
    }
    if (!Y && !X) {  // header only
      ok = true;
      break;
    }
    if (!Y || !X || ldy < m->num_data || ldx < m->num_data) {
      r.fail("Y and X buffers (ld >= numData) are both required");
      break;
    }
    const int d = m->output_dim, q = m->latent_dim;
    bool rows_ok = true;
    for (int64_t i = 0; i < m->num_data && rows_ok; i++) {
      if (!r.line(l)) {
        rows_ok = r.fail("end of file inside the data rows");
        break;
      }
      std::vector<std::string> tok;
      size_t last = l.find_first_not_of(' ', 0), p2 = l.find_first_of(' ', last);
      while (p2 != std::string::npos || last != std::string::npos) {
        tok.push_back(l.substr(last, p2 - last));
        last = l.find_first_not_of(' ', p2);
        p2 = l.find_first_of(' ', last);
      }
      if ((int)tok.size() < d + q + (m->has_labels ? 1 : 0)) {  // the reference indexes past the end here
        rows_ok = r.fail("too few values in a data row");
        break;
      }
      for (int j = 0; j < d; j++) Y[i + (int64_t)j * ldy] = atof(tok[(size_t)j].c_str());
      for (int j = 0; j < q; j++) X[i + (int64_t)j * ldx] = atof(tok[(size_t)(d + j)].c_str());
      if (m->has_labels && labels) labels[i] = atoi(tok[(size_t)(d + q)].c_str());
    }
    ok = rows_ok;
  } while (false);
  if (!ok) {
    std::string e = r.err.empty() ? std::string("gpc_gplvm_model_read: failed on ") + path : r.err;
    size_t at = e.find("gpc_gp_model_read");
    if (at != std::string::npos) e.replace(at, strlen("gpc_gp_model_read"), "gpc_gplvm_model_read");
    set_error(e);
    return GPC_ERR_ARG;
  }
  return GPC_OK;
}

int gpc_gplvm_model_write(const char* path, const gpc_gplvm_model* m, const double* Y, int64_t ldy, const double* X,
                          int64_t ldx, const int* labels, const char* comment) {
  if (!path || !m || !Y || !X) {
    set_error("gpc_gplvm_model_write: null argument");
    return GPC_ERR_ARG;
  }
  if (m->back_constrained || m->dynamics_learnt || m->num_data < 0 || m->output_dim < 1 || m->output_dim > GPC_MODEL_MAX_OUT ||
      m->latent_dim < 1 || ldy < m->num_data || ldx < m->num_data || (m->has_labels && !labels)) {
    set_error("gpc_gplvm_model_write: unsupported model or sizes out of range");
    return GPC_ERR_ARG;
  }
  int rc = check_kern_noise(&m->kern, &m->noise, "gpc_gplvm_model_write");
  if (rc != GPC_OK) return rc;
  FILE* f = fopen(path, "wb");
  if (!f) {
    set_error(std::string("gpc_gplvm_model_write: cannot open ") + path);
    return GPC_ERR_ARG;
  }
  // writeGplvmToFile sets ios::scientific BEFORE toStream adds ios::fixed (CGplvm.cpp:908-921): here even the outermost
  // version is hexfloat; the data rows are plain `out << value`, i.e. "%a" for every entry, integers included
  if (comment && comment[0]) fprintf(f, "# %s\n", comment);
  fputs(NESTED_VERSION, f);
  fprintf(f, "baseType=dataModel\ntype=gplvm\nnumData=%lld\noutputDim=%d\ninputDim=%d\n", (long long)m->num_data, m->output_dim,
          m->latent_dim);
  fprintf(f, "latentRegularised=%d\nbackConstrained=0\ndynamicsLearnt=0\n", m->latent_regularised ? 1 : 0);
  put_kern(f, &m->kern);
  put_noise(f, &m->noise);
  fprintf(f, "Y:%d,X:%d%s\n", m->output_dim, m->latent_dim, m->has_labels ? ",labels:1" : "");
  for (int64_t i = 0; i < m->num_data; i++) {
    for (int j = 0; j < m->output_dim; j++) fprintf(f, "%a ", Y[i + (int64_t)j * ldy]);
    for (int j = 0; j < m->latent_dim; j++) fprintf(f, "%a ", X[i + (int64_t)j * ldx]);
    if (m->has_labels) fprintf(f, "%d", labels[i]);
    fputc('\n', f);
  }
  const bool bad = ferror(f) != 0;
  if (fclose(f) != 0 || bad) {
    set_error(std::string("gpc_gplvm_model_write: write error on ") + path);
    return GPC_ERR_ARG;
  }
  return GPC_OK;
}
