// onnx_import.cu -- ONNX ModelProto -> OARG layer list, host code only (no CUDA calls).
//
// The reference hands ONNX bytes or a path to ONNX Runtime:
//   OrtInfer::new / from_config(ModelSource::{Path, Memory})   oar-ocr-core/src/core/inference/ort_infer_builders.rs:9-70
//   ModelSource                                                oar-ocr-core/src/core/config/model_source.rs:20-28
// This library executes OARG layer lists (oar_ocr_b200/models.py), so a drop-in for that boundary has to accept the
// same bytes behind the C ABI: oar_model_load_onnx (capi.cu) = oar_onnx_to_oarg (here) + oar_model_load_blob.
//
// The operator subset, the pattern matching and the order in which ops and tensors are numbered are those of the
// Python importer (oar_ocr_b200/onnx_io.py: import_onnx), which remains the offline tool; the two are held together by
// tests/test_onnx_cabi.py (byte-identical blobs for every exported graph).  Protobuf is read by hand: ModelProto.graph
// (7) -> GraphProto{node 1, initializer 5, input 11, output 12}, NodeProto{input 1, output 2, name 3, op_type 4,
// attribute 5}, AttributeProto{name 1, f 2, i 3, s 4, floats 7, ints 8, type 20}, TensorProto{dims 1, data_type 2,
// float_data 4, int64_data 7, name 8, raw_data 9}.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "engine.cuh"

namespace oar {
namespace {

[[noreturn]] void bad(const std::string& msg) {
  set_error("ONNX import: %s", msg.c_str());
  throw OarError{OAR_E_MODEL};
}

// ---------------------------------------------------------------------------------------------------------------
// protobuf wire format
// ---------------------------------------------------------------------------------------------------------------
struct PbField {
  int wire = 0;
  uint64_t v = 0;              // varint, fixed32 or fixed64 payload
  const uint8_t* p = nullptr;  // length-delimited payload
  size_t n = 0;
};
using PbMsg = std::map<int, std::vector<PbField>>;

bool read_varint(const uint8_t* buf, size_t len, size_t& pos, uint64_t& out) {
  uint64_t v = 0;
  for (int shift = 0; shift < 64; shift += 7) {
    if (pos >= len) return false;
    const uint8_t b = buf[pos++];
    v |= (uint64_t)(b & 0x7F) << shift;
    if (!(b & 0x80)) {
      out = v;
      return true;
    }
  }
  return false;
}

PbMsg parse_message(const uint8_t* buf, size_t len) {
  PbMsg out;
  size_t pos = 0;
  while (pos < len) {
    uint64_t key;
    if (!read_varint(buf, len, pos, key)) bad("truncated protobuf key");
    PbField f;
    f.wire = (int)(key & 7);
    const int field = (int)(key >> 3);
    if (f.wire == 0) {
      if (!read_varint(buf, len, pos, f.v)) bad("truncated protobuf varint");
    } else if (f.wire == 1) {
      if (len - pos < 8) bad("truncated protobuf fixed64");
      memcpy(&f.v, buf + pos, 8);
      pos += 8;
    } else if (f.wire == 2) {
      uint64_t n;
      if (!read_varint(buf, len, pos, n) || n > len - pos) bad("truncated protobuf bytes field");
      f.p = buf + pos, f.n = (size_t)n;
      pos += (size_t)n;
    } else if (f.wire == 5) {
      if (len - pos < 4) bad("truncated protobuf fixed32");
      uint32_t w;
      memcpy(&w, buf + pos, 4);
      f.v = w;
      pos += 4;
    } else {
      bad("unsupported protobuf wire type " + std::to_string(f.wire));
    }
    out[field].push_back(f);
  }
  return out;
}

const std::vector<PbField>& fields(const PbMsg& m, int id) {
  static const std::vector<PbField> none;
  auto it = m.find(id);
  return it == m.end() ? none : it->second;
}
std::string str_of(const PbField& f) { return std::string((const char*)f.p, f.n); }

// repeated int64: packed (one bytes blob per field) or one varint per field
std::vector<int64_t> ints_of(const std::vector<PbField>& fs) {
  std::vector<int64_t> out;
  for (const PbField& f : fs) {
    if (f.wire == 0) {
      out.push_back((int64_t)f.v);
    } else if (f.wire == 2) {
      size_t pos = 0;
      while (pos < f.n) {
        uint64_t v;
        if (!read_varint(f.p, f.n, pos, v)) bad("truncated packed integers");
        out.push_back((int64_t)v);
      }
    }
  }
  return out;
}
std::vector<float> floats_of(const std::vector<PbField>& fs) {
  std::vector<float> out;
  for (const PbField& f : fs) {
    if (f.wire == 5) {
      uint32_t w = (uint32_t)f.v;
      float x;
      memcpy(&x, &w, 4);
      out.push_back(x);
    } else if (f.wire == 2) {
      for (size_t i = 0; i + 4 <= f.n; i += 4) {
        float x;
        memcpy(&x, f.p + i, 4);
        out.push_back(x);
      }
    }
  }
  return out;
}

// ---------------------------------------------------------------------------------------------------------------
// ONNX graph in memory
// ---------------------------------------------------------------------------------------------------------------
constexpr int DT_FLOAT = 1, DT_INT64 = 7;

struct Init {
  std::vector<int64_t> dims;
  int dtype = DT_FLOAT;
  std::vector<float> f;
  std::vector<int64_t> i;
  size_t size() const { return dtype == DT_FLOAT ? f.size() : i.size(); }
  int64_t dim(size_t k) const { return k < dims.size() ? dims[k] : 1; }
  double scalar() const { return dtype == DT_FLOAT ? (double)f.at(0) : (double)i.at(0); }
};

struct Attr {
  bool has_f = false, has_i = false, has_s = false, has_ints = false, has_floats = false;
  float f = 0.0f;
  int64_t i = 0;
  std::string s;
  std::vector<int64_t> ints;
  std::vector<float> floats;
};

struct Node {
  std::string op, name;
  std::vector<std::string> in, out;
  std::map<std::string, Attr> attrs;
  const Attr* attr(const char* k) const {
    auto it = attrs.find(k);
    return it == attrs.end() ? nullptr : &it->second;
  }
  int64_t geti(const char* k, int64_t d) const {
    const Attr* a = attr(k);
    return a && a->has_i ? a->i : (a ? 0 : d);  // a present scalar with its zero value omitted reads as 0
  }
  float getf(const char* k, float d) const {
    const Attr* a = attr(k);
    return a && a->has_f ? a->f : (a ? 0.0f : d);
  }
  std::string gets(const char* k, const char* d) const {
    const Attr* a = attr(k);
    return a ? a->s : std::string(d);
  }
  std::vector<int64_t> getints(const char* k, std::vector<int64_t> d) const {
    const Attr* a = attr(k);
    return a ? a->ints : d;
  }
};

struct Onnx {
  std::vector<Node> nodes;
  std::map<std::string, Init> inits;
  std::vector<std::string> inputs, outputs;  // graph inputs that are not initializers; graph outputs
};

Onnx read_model(const uint8_t* data, size_t len) {
  PbMsg model = parse_message(data, len);
  if (fields(model, 7).empty()) bad("ONNX model has no graph");
  const PbField& gf = fields(model, 7)[0];
  if (gf.wire != 2) bad("ONNX model has no graph");
  PbMsg g = parse_message(gf.p, gf.n);
  Onnx o;
  for (const PbField& raw : fields(g, 5)) {
    PbMsg t = parse_message(raw.p, raw.n);
    Init in;
    in.dims = ints_of(fields(t, 1));
    in.dtype = fields(t, 2).empty() ? DT_FLOAT : (int)fields(t, 2)[0].v;
    std::string name = fields(t, 8).empty() ? std::string() : str_of(fields(t, 8)[0]);
    if (in.dtype != DT_FLOAT && in.dtype != DT_INT64)
      bad("initializer " + name + ": unsupported data type " + std::to_string(in.dtype));
    if (!fields(t, 9).empty()) {
      const PbField& r = fields(t, 9)[0];
      if (in.dtype == DT_FLOAT) {
        in.f.resize(r.n / 4);
        memcpy(in.f.data(), r.p, in.f.size() * 4);
      } else {
        in.i.resize(r.n / 8);
        memcpy(in.i.data(), r.p, in.i.size() * 8);
      }
    } else if (in.dtype == DT_FLOAT) {
      in.f = floats_of(fields(t, 4));
    } else {
      in.i = ints_of(fields(t, 7));
    }
    int64_t want = 1;
    for (int64_t d : in.dims) {
      if (d < 0 || (d > 0 && want > (int64_t)1 << 40)) bad("initializer " + name + ": implausible shape");
      want *= d;
    }
    if ((int64_t)in.size() != want) bad("initializer " + name + ": data does not match its shape");
    o.inits[name] = std::move(in);
  }
  for (const PbField& raw : fields(g, 1)) {
    PbMsg n = parse_message(raw.p, raw.n);
    Node nd;
    if (fields(n, 4).empty()) bad("node without op_type");
    nd.op = str_of(fields(n, 4)[0]);
    nd.name = fields(n, 3).empty() ? std::string() : str_of(fields(n, 3)[0]);
    for (const PbField& f : fields(n, 1)) nd.in.push_back(str_of(f));
    for (const PbField& f : fields(n, 2)) nd.out.push_back(str_of(f));
    for (const PbField& ra : fields(n, 5)) {
      PbMsg a = parse_message(ra.p, ra.n);
      if (fields(a, 1).empty()) continue;
      Attr at;
      if (!fields(a, 2).empty()) {
        uint32_t w = (uint32_t)fields(a, 2)[0].v;
        memcpy(&at.f, &w, 4);
        at.has_f = true;
      }
      if (!fields(a, 3).empty()) at.i = (int64_t)fields(a, 3)[0].v, at.has_i = true;
      if (!fields(a, 4).empty()) at.s = str_of(fields(a, 4)[0]), at.has_s = true;
      if (!fields(a, 8).empty()) at.ints = ints_of(fields(a, 8)), at.has_ints = true;
      if (!fields(a, 7).empty()) at.floats = floats_of(fields(a, 7)), at.has_floats = true;
      nd.attrs[str_of(fields(a, 1)[0])] = std::move(at);
    }
    if (nd.out.empty()) bad("node '" + nd.name + "' has no output");
    o.nodes.push_back(std::move(nd));
  }
  auto names = [&](int field) {
    std::vector<std::string> out;
    for (const PbField& v : fields(g, field)) {
      PbMsg vi = parse_message(v.p, v.n);
      if (!fields(vi, 1).empty()) out.push_back(str_of(fields(vi, 1)[0]));
    }
    return out;
  };
  for (const std::string& n : names(11))
    if (!o.inits.count(n)) o.inputs.push_back(n);
  o.outputs = names(12);
  return o;
}

// ---------------------------------------------------------------------------------------------------------------
// OARG graph under construction (models.py: GraphBuilder)
// ---------------------------------------------------------------------------------------------------------------
constexpr int OPX_PAD = 12, OPX_MAXPOOL = 13;  // carried by the format; the CUDA engine rejects them at load time

struct GOp {
  int type = 0, in0 = 0, in1 = -1, out = 0;
  int32_t p[12] = {0};
  float f[4] = {0, 0, 0, 0};
  std::vector<float> w[4];
  int nw = 0;
};

struct Graph {
  std::vector<GOp> ops;
  std::vector<int> channels{3};  // tensor 0 = the network input, 3 channels
  int new_tensor(int c) {
    channels.push_back(c);
    return (int)channels.size() - 1;
  }
  GOp& push(int type, int in0, int in1, int out) {
    ops.emplace_back();
    GOp& o = ops.back();
    o.type = type, o.in0 = in0, o.in1 = in1, o.out = out;
    return o;
  }
  int conv(int x, int cout, int kh, int kw, int sh, int sw, int ph, int pw, int act, float ps, float pb,
           std::vector<float> w, std::vector<float> b) {
    const int cin = channels[x];
    const int out = new_tensor(cout);
    GOp& o = push(OP_CONV, x, -1, out);
    const int32_t p[12] = {kh, kw, sh, sw, ph, pw, cin, cout, act, 1, 0, 0};
    memcpy(o.p, p, sizeof(p));
    o.f[0] = ps, o.f[1] = pb;
    o.w[0] = std::move(w), o.w[1] = std::move(b), o.nw = 2;
    return out;
  }
  int dwconv(int x, int k, int sh, int sw, int act, float ps, float pb, std::vector<float> w, std::vector<float> b) {
    const int c = channels[x];
    const int out = new_tensor(c);
    GOp& o = push(OP_DWCONV, x, -1, out);
    const int32_t p[12] = {k, k, sh, sw, k / 2, k / 2, c, act, 1, 0, 0, 0};
    memcpy(o.p, p, sizeof(p));
    o.f[0] = ps, o.f[1] = pb;
    o.w[0] = std::move(w), o.w[1] = std::move(b), o.nw = 2;
    return out;
  }
  int add(int a, int b) {
    const int out = new_tensor(channels[a]);
    push(OP_ADD, a, b, out);
    return out;
  }
  int upadd(int a, int b, int scale) {
    const int out = new_tensor(channels[a]);
    push(OP_UPADD, a, b, out).p[0] = scale;
    return out;
  }
  void upsample_into(int x, int scale, int out, int c_off, int c_total) {
    GOp& o = push(OP_UPSAMPLE, x, -1, out);
    o.p[0] = scale, o.p[10] = c_off, o.p[11] = c_total;
  }
  int deconv2(int x, int cout, int act, std::vector<float> w, std::vector<float> b) {
    const int cin = channels[x];
    const int out = new_tensor(cout);
    GOp& o = push(OP_DECONV2, x, -1, out);
    o.p[0] = cin, o.p[1] = cout, o.p[2] = act;
    o.f[0] = 1.0f;
    o.w[0] = std::move(w), o.w[1] = std::move(b), o.nw = 2;
    return out;
  }
  int pool(int type, int x, int k0, int k1, int s0, int s1) {
    const int out = new_tensor(channels[x]);
    GOp& o = push(type, x, -1, out);
    o.p[0] = k0, o.p[1] = k1, o.p[2] = s0, o.p[3] = s1;
    return out;
  }
  int ctc_head(int x, int vocab, std::vector<float> w, std::vector<float> b) {
    const int c = channels[x];
    const int out = new_tensor(vocab);
    GOp& o = push(OP_CTC_HEAD, x, -1, out);
    o.p[0] = c, o.p[1] = vocab;
    o.w[0] = std::move(w), o.w[1] = std::move(b), o.nw = 2;
    return out;
  }
  std::vector<uint8_t> serialize(int kind) const {
    std::vector<OpRec> recs(ops.size());
    std::vector<float> weights;
    for (size_t i = 0; i < ops.size(); ++i) {
      const GOp& o = ops[i];
      OpRec& r = recs[i];
      memset(&r, 0, sizeof(r));
      r.type = o.type, r.in0 = o.in0, r.in1 = o.in1, r.out = o.out;
      memcpy(r.p, o.p, sizeof(r.p));
      memcpy(r.f, o.f, sizeof(r.f));
      for (int k = 0; k < o.nw; ++k) {
        r.w_off[k] = (int64_t)weights.size();
        r.w_len[k] = (int64_t)o.w[k].size();
        weights.insert(weights.end(), o.w[k].begin(), o.w[k].end());
        while (weights.size() & 3) weights.push_back(0.0f);  // every array stays 16-byte aligned
      }
    }
    std::vector<uint8_t> blob(28 + recs.size() * sizeof(OpRec) + weights.size() * 4);
    memcpy(blob.data(), "OARG", 4);
    const uint32_t head[4] = {1u, (uint32_t)kind, (uint32_t)ops.size(), (uint32_t)channels.size()};
    memcpy(blob.data() + 4, head, 16);
    const uint64_t nw = weights.size();
    memcpy(blob.data() + 20, &nw, 8);
    if (!recs.empty()) memcpy(blob.data() + 28, recs.data(), recs.size() * sizeof(OpRec));
    if (nw) memcpy(blob.data() + 28 + recs.size() * sizeof(OpRec), weights.data(), nw * 4);
    return blob;
  }
};

// ---------------------------------------------------------------------------------------------------------------
// the importer
// ---------------------------------------------------------------------------------------------------------------
struct Importer {
  const Onnx& m;
  Graph g;
  std::map<std::string, std::vector<int>> consumers;  // value -> consuming node indices (-1 = graph output)
  std::map<std::string, int> tid;                     // ONNX value -> OARG tensor id
  std::set<int> done;
  int kind = OAR_KIND_DET;

  explicit Importer(const Onnx& model) : m(model) {}

  [[noreturn]] void fail(const Node& n, const std::string& why) const {
    bad("node '" + n.name + "' (" + n.op + "): " + why);
  }
  const std::vector<int>& cons(const std::string& v) const {
    static const std::vector<int> none;
    auto it = consumers.find(v);
    return it == consumers.end() ? none : it->second;
  }
  const Init* init(const std::string& name) const {
    auto it = m.inits.find(name);
    return it == m.inits.end() ? nullptr : &it->second;
  }
  const Init& need_init(const Node& n, const std::string& name, const char* what) const {
    const Init* p = init(name);
    if (!p) fail(n, std::string(what) + " is not an initializer");
    return *p;
  }
  std::vector<float> as_f32(const Init& t) const {
    if (t.dtype == DT_FLOAT) return t.f;
    std::vector<float> out(t.i.size());
    for (size_t k = 0; k < out.size(); ++k) out[k] = (float)t.i[k];
    return out;
  }
  int tensor_of(const Node& n, const std::string& v) const {
    auto it = tid.find(v);
    if (it == tid.end())
      fail(n, "operand '" + v + "' is " +
                  (init(v) ? "an initializer where an activation is expected" : "not produced by a supported node"));
    return it->second;
  }
  // the single not-yet-consumed consumer of `name` (optionally of a given type), or -1
  int sole_consumer(const std::string& name, const char* op = nullptr) const {
    const std::vector<int>& c = cons(name);
    if (c.size() != 1 || c[0] < 0 || done.count(c[0])) return -1;
    return (!op || m.nodes[c[0]].op == op) ? c[0] : -1;
  }
  bool feeds_mul_with(const Node& n, const std::string& x) const {
    const int i = sole_consumer(n.out[0], "Mul");
    return i >= 0 && std::count(m.nodes[i].in.begin(), m.nodes[i].in.end(), x) > 0;
  }
  static bool near(float a, float b) { return std::fabs((double)a - (double)b) <= 1e-6; }

  struct ActOut {
    int act = ACT_NONE;
    float ps = 1.0f, pb = 0.0f;
    std::string y;
  };
  // consume the activation (and learnable affine) that follows value y
  ActOut take_activation(std::string y) {
    ActOut r;
    int i = sole_consumer(y);
    if (i >= 0) {
      const Node& n = m.nodes[i];
      const int simple = n.op == "Relu" ? ACT_RELU : n.op == "HardSwish" ? ACT_HSWISH : n.op == "Sigmoid" ? ACT_SIGMOID : -1;
      if (simple >= 0 && !(n.op == "Sigmoid" && feeds_mul_with(n, y))) {
        r.act = simple, y = n.out[0];
        done.insert(i);
      } else if (n.op == "HardSigmoid" && !feeds_mul_with(n, y)) {
        if (!near(n.getf("alpha", 0.2f), 1.0f / 6.0f) || !near(n.getf("beta", 0.5f), 0.5f))
          fail(n, "HardSigmoid activation with alpha/beta other than 1/6, 0.5");
        r.act = ACT_HSIGMOID, y = n.out[0];
        done.insert(i);
      }
    }
    if (r.act == ACT_NONE) {  // decomposed forms: x * Sigmoid(x) (swish), x * HardSigmoid(x; 1/6, 0.5) (hardswish)
      std::vector<int> c;
      for (int k : cons(y))
        if (k >= 0 && !done.count(k)) c.push_back(k);
      if (c.size() == 2) {
        std::vector<int> gate, mul;
        for (int k : c) {
          if (m.nodes[k].op == "Sigmoid" || m.nodes[k].op == "HardSigmoid") gate.push_back(k);
          if (m.nodes[k].op == "Mul") mul.push_back(k);
        }
        if (gate.size() == 1 && mul.size() == 1) {
          const Node& gn = m.nodes[gate[0]];
          const Node& mn = m.nodes[mul[0]];
          std::set<std::string> have(mn.in.begin(), mn.in.end()), want{y, gn.out[0]};
          if (have == want) {
            if (gn.op == "Sigmoid") {
              r.act = ACT_SWISH;
            } else {
              if (!near(gn.getf("alpha", 0.2f), 1.0f / 6.0f) || !near(gn.getf("beta", 0.5f), 0.5f))
                fail(gn, "decomposed hardswish with alpha/beta other than 1/6, 0.5");
              r.act = ACT_HSWISH;
            }
            done.insert(gate[0]);
            done.insert(mul[0]);
            y = mn.out[0];
          }
        }
      }
    }
    // learnable affine: Mul by a scalar initializer then Add of a scalar initializer
    i = sole_consumer(y, "Mul");
    if (i >= 0 && r.act != ACT_NONE) {
      std::vector<std::string> other;
      for (const std::string& x : m.nodes[i].in)
        if (x != y) other.push_back(x);
      const Init* a = other.size() == 1 ? init(other[0]) : nullptr;
      if (a && a->size() == 1) {
        const std::string& my = m.nodes[i].out[0];
        const int j = sole_consumer(my, "Add");
        if (j >= 0) {
          std::vector<std::string> ob;
          for (const std::string& x : m.nodes[j].in)
            if (x != my) ob.push_back(x);
          const Init* b = ob.size() == 1 ? init(ob[0]) : nullptr;
          if (b && b->size() == 1) {
            r.ps = (float)a->scalar(), r.pb = (float)b->scalar();
            done.insert(i);
            done.insert(j);
            y = m.nodes[j].out[0];
          }
        }
      }
    }
    r.y = y;
    return r;
  }

  struct ConvW {
    std::vector<float> w, b;  // w in ONNX layout [d0][d1][kh][kw]
    int d0 = 0, d1 = 0, kh = 0, kw = 0, sh = 1, sw = 1, ph = 0, pw = 0, group = 1;
  };
  ConvW conv_params(const Node& n) {
    if (n.in.size() < 2) fail(n, "convolution without weights");
    const Init& w = need_init(n, n.in[1], "weights");
    if (w.dims.size() != 4) fail(n, "weights are not 4-D");
    ConvW c;
    c.d0 = (int)w.dims[0], c.d1 = (int)w.dims[1];
    std::vector<int64_t> ks = n.getints("kernel_shape", {w.dims[2], w.dims[3]});
    std::vector<int64_t> st = n.getints("strides", {1, 1});
    std::vector<int64_t> pads = n.getints("pads", {0, 0, 0, 0});
    if (ks.size() != 2 || st.size() != 2 || pads.size() != 4) fail(n, "only 2-D convolutions are supported");
    c.kh = (int)ks[0], c.kw = (int)ks[1], c.sh = (int)st[0], c.sw = (int)st[1];
    if (c.kh != w.dims[2] || c.kw != w.dims[3]) fail(n, "kernel_shape does not match the weights");
    std::string auto_pad = n.gets("auto_pad", "NOTSET");
    if (auto_pad.empty()) auto_pad = "NOTSET";
    if (auto_pad == "VALID") {
      pads = {0, 0, 0, 0};
    } else if (auto_pad == "SAME_UPPER" || auto_pad == "SAME_LOWER") {
      if (c.sh != 1 || c.sw != 1 || c.kh % 2 == 0 || c.kw % 2 == 0)
        fail(n, "auto_pad=" + auto_pad + " with stride " + std::to_string(c.sh) + "x" + std::to_string(c.sw) + " / kernel " +
                    std::to_string(c.kh) + "x" + std::to_string(c.kw) + " (needs stride 1 and an odd kernel)");
      pads = {c.kh / 2, c.kw / 2, c.kh / 2, c.kw / 2};
    } else if (auto_pad != "NOTSET") {
      fail(n, "unknown auto_pad '" + auto_pad + "'");
    }
    if (pads[0] != pads[2] || pads[1] != pads[3]) fail(n, "asymmetric padding");
    for (int64_t d : n.getints("dilations", {1, 1}))
      if (d != 1) fail(n, "dilation");
    c.ph = (int)pads[0], c.pw = (int)pads[1];
    c.group = (int)n.geti("group", 1);
    c.w = as_f32(w);
    if (n.in.size() > 2 && !n.in[2].empty())
      c.b = as_f32(need_init(n, n.in[2], "bias"));
    else
      c.b.assign((size_t)c.d0, 0.0f);
    return c;
  }

  // BatchNormalization folded into the producer: k = scale / sqrt(var + eps) per output channel (channel = dim `axis`
  // of a [d0][d1][kh][kw] weight), w *= k, b = (b - mean) * k + beta -- in f32, operation by operation, as numpy does
  std::string fold_bn(const std::string& y, ConvW& c, int axis) {
    const int i = sole_consumer(y, "BatchNormalization");
    if (i < 0) return y;
    const Node& n = m.nodes[i];
    if (n.in.size() < 5) fail(n, "BatchNormalization needs scale, bias, mean and variance");
    std::vector<float> sc = as_f32(need_init(n, n.in[1], "scale")), bi = as_f32(need_init(n, n.in[2], "bias")),
                       mean = as_f32(need_init(n, n.in[3], "mean")), var = as_f32(need_init(n, n.in[4], "variance"));
    const size_t C = axis == 0 ? c.d0 : c.d1;
    if (sc.size() != C || bi.size() != C || mean.size() != C || var.size() != C || c.b.size() != C)
      fail(n, "BatchNormalization parameters do not match the channel count");
    const float eps = n.getf("epsilon", 1e-5f);
    std::vector<float> k(C);
    for (size_t ch = 0; ch < C; ++ch) k[ch] = sc[ch] / std::sqrt(var[ch] + eps);
    const size_t inner = (size_t)c.kh * c.kw;
    for (int a = 0; a < c.d0; ++a)
      for (int b = 0; b < c.d1; ++b) {
        const float kk = k[axis == 0 ? a : b];
        float* wp = c.w.data() + ((size_t)a * c.d1 + b) * inner;
        for (size_t e = 0; e < inner; ++e) wp[e] = wp[e] * kk;
      }
    for (size_t ch = 0; ch < C; ++ch) {
      const float t = (c.b[ch] - mean[ch]) * k[ch];
      c.b[ch] = t + bi[ch];
    }
    done.insert(i);
    return n.out[0];
  }

  // follow single-consumer links from `start` through the given op types
  std::vector<int> chain(const std::string& start, std::initializer_list<const char*> types, const Node& ctx) {
    std::vector<int> seq;
    std::string cur = start;
    for (const char* want : types) {
      std::vector<int> c;
      for (int k : cons(cur))
        if (k >= 0 && !done.count(k) && m.nodes[k].op == want) c.push_back(k);
      if (c.size() != 1) fail(ctx, std::string("expected ") + want + " after '" + cur + "'");
      seq.push_back(c[0]);
      cur = m.nodes[c[0]].out[0];
    }
    return seq;
  }
  const Init& init_operand(const Node& n) const {  // the one initializer among a node's inputs
    for (const std::string& x : n.in)
      if (const Init* p = init(x)) return *p;
    fail(n, "expected an initializer operand");
  }
  static std::vector<float> transpose2(const std::vector<float>& a, size_t rows, size_t cols) {  // [rows][cols] -> [cols][rows]
    std::vector<float> t(a.size());
    for (size_t r = 0; r < rows; ++r)
      for (size_t c = 0; c < cols; ++c) t[c * rows + r] = a[r * cols + c];
    return t;
  }

  void import_attention(int i) {
    const Node& n = m.nodes[i];
    const std::string& t0 = n.out[0];
    std::vector<int> shape_i;
    for (int k : cons(t0))
      if (k >= 0 && m.nodes[k].op == "Shape") shape_i.push_back(k);
    if (shape_i.size() != 1) fail(n, "attention: missing Shape of the NHWC tensor");
    std::vector<int> seq = chain(t0, {"Reshape", "MatMul", "Add", "Reshape", "Transpose", "Split"}, n);
    for (int k : cons(m.nodes[seq[0]].out[0]))
      if (k >= 0 && m.nodes[k].op == "Shape") shape_i.push_back(k);
    const Node& split = m.nodes[seq.back()];
    if (split.out.size() != 3) fail(split, "attention: Split must produce q, k and v");
    std::vector<int> qs = chain(split.out[0], {"Mul"}, n);
    std::vector<int> kt = chain(split.out[1], {"Transpose"}, n);
    std::vector<int> tail = chain(m.nodes[qs[0]].out[0],
                                  {"MatMul", "Softmax", "MatMul", "Transpose", "Reshape", "MatMul", "Add", "Reshape", "Transpose"}, n);
    const Init& wqkv = need_init(m.nodes[seq[1]], m.nodes[seq[1]].in.at(1), "qkv weights");
    const Init& bqkv = init_operand(m.nodes[seq[2]]);
    const Init& shp5 = need_init(m.nodes[seq[3]], m.nodes[seq[3]].in.at(1), "reshape target");
    if (shp5.dtype != DT_INT64 || shp5.i.size() != 5) fail(n, "attention: head reshape is not [0,0,3,heads,d]");
    const int heads = (int)shp5.i[3];
    const float scale = (float)init_operand(m.nodes[qs[0]]).scalar();
    const Init& wp = need_init(m.nodes[tail[5]], m.nodes[tail[5]].in.at(1), "projection weights");
    const Init& bp = init_operand(m.nodes[tail[6]]);
    for (auto* v : {&shape_i, &seq, &qs, &kt, &tail})
      for (int k : *v) done.insert(k);
    const int x = tensor_of(n, n.in[0]);
    const int c = g.channels[x];
    if (wqkv.dims.size() != 2 || wqkv.dims[0] != c || wqkv.dims[1] != 3 * c || wp.dims.size() != 2 || wp.dims[0] != c ||
        wp.dims[1] != c || heads <= 0 || c % heads)
      fail(n, "attention: weight shapes do not match the channel count");
    const int out = g.new_tensor(c);
    GOp& o = g.push(OP_ATTN, x, -1, out);
    o.p[0] = c, o.p[1] = heads;
    o.f[0] = scale;
    o.w[0] = transpose2(as_f32(wqkv), c, 3 * (size_t)c), o.w[1] = as_f32(bqkv);
    o.w[2] = transpose2(as_f32(wp), c, c), o.w[3] = as_f32(bp);
    o.nw = 4;
    tid[m.nodes[tail.back()].out[0]] = out;
  }

  void import_ctc_head(int i) {
    const Node& n = m.nodes[i];
    // graph outputs are not consumers here (the Softmax is normally the graph output)
    std::vector<int> seq = chain(n.out[0], {"Reshape", "MatMul", "Add", "Softmax"}, n);
    const Init& w = need_init(m.nodes[seq[1]], m.nodes[seq[1]].in.at(1), "head weights");
    const Init& b = init_operand(m.nodes[seq[2]]);
    for (int k : seq) done.insert(k);
    const int x = tensor_of(n, n.in[0]);
    if (w.dims.size() != 2 || w.dims[0] != g.channels[x]) fail(n, "CTC head weights do not match the feature width");
    const int vocab = (int)w.dims[1];
    tid[m.nodes[seq.back()].out[0]] = g.ctc_head(x, vocab, transpose2(as_f32(w), (size_t)w.dims[0], vocab), as_f32(b));
  }

  void import_cls_head(int i) {
    const Node& n = m.nodes[i];
    auto sole = [&](const std::string& v, std::initializer_list<const char*> types) {
      std::vector<int> c;
      for (int k : cons(v))
        if (k >= 0 && !done.count(k)) c.push_back(k);
      if (c.size() != 1) return -1;
      for (const char* t : types)
        if (m.nodes[c[0]].op == t) return c[0];
      return -1;
    };
    const int x = tensor_of(n, n.in[0]);
    const int j = sole(n.out[0], {"Gemm", "MatMul"});
    if (j < 0) fail(n, "flattened features must feed exactly one Gemm / MatMul");
    const Node& mm = m.nodes[j];
    if (mm.in.size() < 2 || !init(mm.in[1])) fail(mm, "classifier weights must be an initializer");
    const Init& wi = *init(mm.in[1]);
    if (wi.dims.size() != 2) fail(mm, "classifier weights are not 2-D");
    std::vector<float> w = as_f32(wi);
    size_t rows = (size_t)wi.dims[0], cols = (size_t)wi.dims[1];
    std::vector<int> seq{j};
    std::string cur = mm.out[0];
    std::vector<float> b;
    bool has_b = false;
    if (mm.op == "Gemm") {
      if (mm.getf("alpha", 1.0f) != 1.0f || mm.getf("beta", 1.0f) != 1.0f || mm.geti("transA", 0))
        fail(mm, "Gemm with alpha/beta != 1 or transA");
      if (!mm.geti("transB", 0)) w = transpose2(w, rows, cols), std::swap(rows, cols);  // -> [classes][C]
      if (mm.in.size() > 2 && !mm.in[2].empty()) b = as_f32(need_init(mm, mm.in[2], "classifier bias")), has_b = true;
    } else {
      w = transpose2(w, rows, cols), std::swap(rows, cols);
      const int k = sole(cur, {"Add"});
      if (k >= 0) {
        int n_init = 0;
        for (const std::string& v : m.nodes[k].in)
          if (init(v)) ++n_init;
        if (n_init != 1) fail(m.nodes[k], "classifier bias must be an initializer");
        b = as_f32(init_operand(m.nodes[k])), has_b = true;
        seq.push_back(k);
        cur = m.nodes[k].out[0];
      }
    }
    const int k = sole(cur, {"Softmax"});
    if (k < 0) fail(mm, "classifier head must end in Softmax");
    const int64_t axis = m.nodes[k].geti("axis", -1);
    if (axis != 1 && axis != -1) fail(m.nodes[k], "Softmax over an axis other than the classes");
    seq.push_back(k);
    if ((int)cols != g.channels[x]) fail(mm, "classifier weights do not match the pooled channels");
    if (!has_b) b.assign(rows, 0.0f);
    for (int q : seq) done.insert(q);
    tid[m.nodes[k].out[0]] = g.ctc_head(x, (int)rows, std::move(w), std::move(b));
  }

  std::vector<uint8_t> run(int seed_kind) {
    if (m.inputs.size() != 1) bad("expected one graph input, found " + std::to_string(m.inputs.size()));
    if (m.outputs.empty()) bad("graph has no output");
    for (size_t i = 0; i < m.nodes.size(); ++i)
      for (const std::string& x : m.nodes[i].in) consumers[x].push_back((int)i);
    for (const std::string& o : m.outputs) consumers[o].push_back(-1);
    tid[m.inputs[0]] = 0;
    for (const Node& n : m.nodes)
      if (n.op == "Concat" && n.geti("axis", 1) != 1) fail(n, "Concat on an axis other than channels");
    for (int i = 0; i < (int)m.nodes.size(); ++i) {
      if (done.count(i)) continue;
      done.insert(i);
      const Node& n = m.nodes[i];
      const std::string& ot = n.op;
      if (n.in.empty()) fail(n, "node without inputs");
      if (ot == "Conv") {
        ConvW c = conv_params(n);
        std::string y = fold_bn(n.out[0], c, 0);
        const int x = tensor_of(n, n.in[0]);
        const int cin = g.channels[x];
        if (c.group == 1) {
          if (c.d1 != cin) fail(n, "weights expect " + std::to_string(c.d1) + " input channels, the tensor has " + std::to_string(cin));
          ActOut a = take_activation(y);
          // [cout][cin][kh][kw] -> [cout][kh][kw][cin]
          std::vector<float> w(c.w.size());
          for (int co = 0; co < c.d0; ++co)
            for (int ci = 0; ci < c.d1; ++ci)
              for (int ky = 0; ky < c.kh; ++ky)
                for (int kx = 0; kx < c.kw; ++kx)
                  w[(((size_t)co * c.kh + ky) * c.kw + kx) * c.d1 + ci] = c.w[(((size_t)co * c.d1 + ci) * c.kh + ky) * c.kw + kx];
          tid[a.y] = g.conv(x, c.d0, c.kh, c.kw, c.sh, c.sw, c.ph, c.pw, a.act, a.ps, a.pb, std::move(w), std::move(c.b));
        } else if (c.group == cin && c.d0 == cin && c.d1 == 1 && c.kh == c.kw && c.ph == c.kh / 2 && c.pw == c.kh / 2) {
          ActOut a = take_activation(y);
          // [c][1][k][k] -> [k][k][c]
          std::vector<float> w(c.w.size());
          const int k = c.kh;
          for (int ch = 0; ch < cin; ++ch)
            for (int e = 0; e < k * k; ++e) w[(size_t)e * cin + ch] = c.w[(size_t)ch * k * k + e];
          tid[a.y] = g.dwconv(x, k, c.sh, c.sw, a.act, a.ps, a.pb, std::move(w), std::move(c.b));
        } else {
          fail(n, "grouped convolution (group=" + std::to_string(c.group) + ") other than depthwise");
        }
      } else if (ot == "GlobalAveragePool") {
        // squeeze-excite: GAP -> Conv1x1 -> Relu -> Conv1x1 -> HardSigmoid -> Mul(x, gate) [-> Add(x, .)];
        // anything else is a plain global pool (the classifier trunk's): AVGPOOL with a (0, 0) window
        std::vector<int> seq;
        std::string cur = n.out[0];
        bool is_se = true;
        for (const char* want : {"Conv", "Relu", "Conv", "HardSigmoid", "Mul"}) {
          const int j = sole_consumer(cur, want);
          if (j < 0) {
            is_se = false;
            break;
          }
          seq.push_back(j);
          cur = m.nodes[j].out[0];
        }
        if (!is_se) {
          tid[n.out[0]] = g.pool(OP_AVGPOOL, tensor_of(n, n.in[0]), 0, 0, 0, 0);
          continue;
        }
        const Node &c1 = m.nodes[seq[0]], &c2 = m.nodes[seq[2]], &hs = m.nodes[seq[3]], &mul = m.nodes[seq[4]];
        if (!std::count(mul.in.begin(), mul.in.end(), n.in[0])) fail(n, "squeeze-excite gate does not multiply the pooled tensor");
        for (int k : seq) done.insert(k);
        auto se_fc = [&](const Node& c, std::vector<float>& w, std::vector<float>& b, int& rows, int& cols) {
          if (c.in.size() < 2) fail(c, "squeeze-excite convolution without weights");
          const Init& wi = need_init(c, c.in[1], "squeeze-excite weights");
          if (wi.dims.size() != 4 || wi.dims[2] != 1 || wi.dims[3] != 1 || c.geti("group", 1) != 1)
            fail(c, "squeeze-excite convolutions must be dense 1x1");
          rows = (int)wi.dims[0], cols = (int)wi.dims[1];
          w = as_f32(wi);
          if (c.in.size() > 2 && !c.in[2].empty())
            b = as_f32(need_init(c, c.in[2], "squeeze-excite bias"));
          else
            b.assign((size_t)rows, 0.0f);
        };
        std::vector<float> w1, b1, w2, b2;
        int r1, k1, r2, k2;
        se_fc(c1, w1, b1, r1, k1);
        se_fc(c2, w2, b2, r2, k2);
        bool residual = false;
        std::string y = mul.out[0];
        const int j = sole_consumer(y, "Add");
        if (j >= 0 && std::count(m.nodes[j].in.begin(), m.nodes[j].in.end(), n.in[0])) {
          residual = true, y = m.nodes[j].out[0];
          done.insert(j);
        }
        const int x = tensor_of(n, n.in[0]);
        const int c = g.channels[x], cm = r1;
        if (k1 != c || r2 != c || k2 != cm) fail(n, "squeeze-excite weights do not match the channel count");
        const int out = g.new_tensor(c);
        GOp& o = g.push(OP_SE, x, -1, out);
        o.p[0] = c, o.p[1] = cm, o.p[2] = residual ? 1 : 0;
        o.f[0] = hs.getf("alpha", 0.2f), o.f[1] = hs.getf("beta", 0.5f);
        o.w[0] = std::move(w1), o.w[1] = std::move(b1), o.w[2] = std::move(w2), o.w[3] = std::move(b2);
        o.nw = 4;
        tid[y] = out;
      } else if (ot == "Add") {
        if (n.in.size() != 2) fail(n, "Add needs two operands");
        const int a = tensor_of(n, n.in[0]), b = tensor_of(n, n.in[1]);
        tid[n.out[0]] = g.add(a, b);
      } else if (ot == "Resize") {
        if (n.gets("mode", "nearest") != "nearest") fail(n, "Resize mode other than nearest");
        const std::string ctm = n.gets("coordinate_transformation_mode", "half_pixel");
        const std::string nm = n.gets("nearest_mode", "round_prefer_floor");
        const bool ok = (ctm == "asymmetric" && nm == "floor") ||
                        ((ctm == "half_pixel" || ctm == "pytorch_half_pixel") &&
                         (nm == "round_prefer_floor" || nm == "round_prefer_ceil"));
        if (!ok)
          fail(n, "Resize nearest with coordinate_transformation_mode=" + ctm + ", nearest_mode=" + nm +
                      " is not integer pixel replication");
        const Init* sc = n.in.size() > 2 && !n.in[2].empty() ? init(n.in[2]) : nullptr;
        if (!sc || sc->dtype != DT_FLOAT || sc->f.size() != 4 || sc->f[0] != 1.0f || sc->f[1] != 1.0f || sc->f[2] != sc->f[3] ||
            sc->f[2] != (float)(int)sc->f[2] || sc->f[2] < 1.0f)
          fail(n, "Resize needs constant integer scales [1,1,s,s]");
        const int s = (int)sc->f[2];
        const int j = sole_consumer(n.out[0], "Add");
        if (j >= 0) {
          std::string other;
          for (const std::string& x : m.nodes[j].in)
            if (x != n.out[0]) {
              other = x;
              break;
            }
          if (other.empty()) fail(m.nodes[j], "Add of an upsampled tensor with itself");
          done.insert(j);
          const int a = tensor_of(m.nodes[j], other), b = tensor_of(n, n.in[0]);
          tid[m.nodes[j].out[0]] = g.upadd(a, b, s);
        } else {
          const int x = tensor_of(n, n.in[0]);
          const int out = g.new_tensor(g.channels[x]);
          g.upsample_into(x, s, out, 0, 0);
          tid[n.out[0]] = out;
        }
      } else if (ot == "Concat") {
        std::vector<int> parts;
        int total = 0;
        for (const std::string& x : n.in) parts.push_back(tensor_of(n, x)), total += g.channels[parts.back()];
        const int out = g.new_tensor(total);
        int off = 0;
        for (int pt : parts) {
          // re-target the producer when it can write a slice itself (conv / upsample that nobody else reads)
          std::vector<size_t> prod;
          for (size_t k = 0; k < g.ops.size(); ++k)
            if (g.ops[k].out == pt) prod.push_back(k);
          std::string name;
          for (const std::string& x : n.in)
            if (tid[x] == pt) {
              name = x;
              break;
            }
          int readers = 0;
          for (const GOp& o : g.ops)
            if (o.in0 == pt || o.in1 == pt) ++readers;
          const bool only_here = cons(name).size() == 1 && readers == 0;
          if (prod.size() == 1 && (g.ops[prod[0]].type == OP_CONV || g.ops[prod[0]].type == OP_UPSAMPLE) && only_here &&
              g.ops[prod[0]].p[11] == 0) {
            GOp& po = g.ops[prod[0]];
            po.out = out, po.p[10] = off, po.p[11] = total;
          } else {
            g.upsample_into(pt, 1, out, off, total);
          }
          off += g.channels[pt];
        }
        tid[n.out[0]] = out;
      } else if (ot == "ConvTranspose") {
        if (n.in.size() < 2) fail(n, "transposed convolution without weights");
        const Init& wi = need_init(n, n.in[1], "weights");  // [cin][cout][kh][kw]
        std::vector<int64_t> st = n.getints("strides", {1, 1}), pads = n.getints("pads", {0, 0, 0, 0});
        bool padded = false;
        for (int64_t p : pads) padded = padded || p != 0;
        if (wi.dims.size() != 4 || wi.dims[2] != 2 || wi.dims[3] != 2 || st != std::vector<int64_t>{2, 2} || padded)
          fail(n, "ConvTranspose other than 2x2 stride 2 without padding");
        ConvW c;
        c.d0 = (int)wi.dims[0], c.d1 = (int)wi.dims[1], c.kh = c.kw = 2;
        c.w = as_f32(wi);
        if (n.in.size() > 2 && !n.in[2].empty())
          c.b = as_f32(need_init(n, n.in[2], "bias"));
        else
          c.b.assign((size_t)c.d1, 0.0f);
        std::string y = fold_bn(n.out[0], c, 1);
        ActOut a = take_activation(y);
        if (a.ps != 1.0f || a.pb != 0.0f) fail(n, "affine after a transposed convolution");
        const int x = tensor_of(n, n.in[0]);
        if (c.d0 != g.channels[x]) fail(n, "weights do not match the input channels");
        // [cin][cout][dy][dx] -> [dy][dx][cout][cin]
        std::vector<float> w(c.w.size());
        for (int ci = 0; ci < c.d0; ++ci)
          for (int co = 0; co < c.d1; ++co)
            for (int q = 0; q < 4; ++q) w[((size_t)q * c.d1 + co) * c.d0 + ci] = c.w[((size_t)ci * c.d1 + co) * 4 + q];
        tid[a.y] = g.deconv2(x, c.d1, a.act, std::move(w), std::move(c.b));
      } else if (ot == "Pad") {
        const std::string mode = n.gets("mode", "constant");
        if (!mode.empty() && mode != "constant") fail(n, "Pad mode other than constant");
        const Init* pi = n.in.size() > 1 ? init(n.in[1]) : nullptr;
        if (!pi || pi->dtype != DT_INT64) fail(n, "Pad amounts must be an initializer");
        const std::vector<int64_t>& pads = pi->i;
        bool okp = pads.size() == 8;
        for (size_t k = 0; okp && k < 8; ++k) okp = pads[k] >= 0 && !((k == 0 || k == 1 || k == 4 || k == 5) && pads[k]);
        if (!okp) fail(n, "Pad must add zeros to the spatial dims of an NCHW tensor");
        if (n.in.size() > 2 && !n.in[2].empty()) {
          const Init* cv = init(n.in[2]);
          if (!cv || cv->size() < 1 || cv->scalar() != 0.0) fail(n, "Pad with a non-zero constant");
        }
        const int x = tensor_of(n, n.in[0]);
        const int out = g.new_tensor(g.channels[x]);
        GOp& o = g.push(OPX_PAD, x, -1, out);
        o.p[0] = (int)pads[2], o.p[1] = (int)pads[3], o.p[2] = (int)pads[6], o.p[3] = (int)pads[7];
        tid[n.out[0]] = out;
      } else if (ot == "MaxPool" || ot == "AveragePool") {
        std::vector<int64_t> ks = n.getints("kernel_shape", {}), pads = n.getints("pads", {0, 0, 0, 0});
        if (ks.size() != 2) fail(n, "pool without a 2-D kernel_shape");
        bool padded = false;
        for (int64_t p : pads) padded = padded || p != 0;
        std::vector<int64_t> st;
        if (ot == "MaxPool") {
          bool dil = false;
          for (int64_t d : n.getints("dilations", {1, 1})) dil = dil || d != 1;
          if (padded || n.geti("ceil_mode", 0) || dil) fail(n, "MaxPool with padding, ceil_mode or dilation");
          st = n.getints("strides", {1, 1});
        } else {
          if (padded) fail(n, "padded AveragePool");
          st = n.getints("strides", ks);
        }
        if (st.size() != 2) fail(n, "pool strides are not 2-D");
        tid[n.out[0]] = g.pool(ot == "MaxPool" ? OPX_MAXPOOL : OP_AVGPOOL, tensor_of(n, n.in[0]), (int)ks[0], (int)ks[1],
                               (int)st[0], (int)st[1]);
      } else if (ot == "Transpose" && n.getints("perm", {}) == std::vector<int64_t>{0, 2, 3, 1}) {
        const int j = sole_consumer(n.out[0]);
        const std::string nxt = j >= 0 ? m.nodes[j].op : std::string();
        if (nxt == "LayerNormalization") {
          const Node& ln = m.nodes[j];
          const int k2 = sole_consumer(ln.out[0], "Transpose");
          if (k2 < 0 || m.nodes[k2].getints("perm", {}) != std::vector<int64_t>{0, 3, 1, 2})
            fail(ln, "LayerNormalization must sit between NHWC/NCHW transposes");
          done.insert(j);
          done.insert(k2);
          const int x = tensor_of(n, n.in[0]);
          const int c = g.channels[x];
          if (ln.in.size() < 3) fail(ln, "LayerNormalization needs scale and bias");
          std::vector<float> gm = as_f32(need_init(ln, ln.in[1], "scale")), be = as_f32(need_init(ln, ln.in[2], "bias"));
          if ((int)gm.size() != c || (int)be.size() != c) fail(ln, "LayerNormalization parameters do not match the channel count");
          const int out = g.new_tensor(c);
          GOp& o = g.push(OP_LAYERNORM, x, -1, out);
          o.p[0] = c;
          o.f[0] = ln.getf("epsilon", 1e-5f);
          o.w[0] = std::move(gm), o.w[1] = std::move(be), o.nw = 2;
          tid[m.nodes[k2].out[0]] = out;
        } else if (cons(n.out[0]).size() == 2) {
          import_attention(i);
        } else if (nxt == "Reshape") {
          kind = OAR_KIND_REC;
          import_ctc_head(i);
        } else {
          fail(n, "NHWC transpose outside LayerNormalization / attention / CTC head");
        }
      } else if (ot == "Dropout" || ot == "Identity") {  // identities at inference time
        tid[n.out[0]] = tensor_of(n, n.in[0]);
      } else if (ot == "Flatten" || ot == "Squeeze" || ot == "Reshape") {
        kind = OAR_KIND_CLS;
        import_cls_head(i);
      } else {
        fail(n, "operator outside the supported subset");
      }
    }
    auto it = tid.find(m.outputs[0]);
    if (it == tid.end()) bad("graph output was not produced");
    if (g.ops.empty() || g.ops.back().out != it->second) bad("the graph output is not the last operation");
    return g.serialize(seed_kind >= 0 ? seed_kind : kind);
  }
};

}  // namespace

std::vector<uint8_t> onnx_to_oarg(const void* bytes, size_t len, int kind_hint) {
  Onnx model = read_model((const uint8_t*)bytes, len);
  Importer imp(model);
  return imp.run(kind_hint);
}

}  // namespace oar
