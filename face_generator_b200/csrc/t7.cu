// Torch7 binary serialisation (the format `torch.save` / `torch.load` use by default) -- host code only.
//
// Why it is here: the reference checkpoints its nets with torch.save(filename, {D=MODEL_D, G=MODEL_G, opt=OPT,
// epoch=EPOCH}) (adversarial.lua:328, adversarial_c2f.lua:216) and sample.lua:251-258 / train.lua:104-124 read them
// back.  A host that is not Torch (the Python mirror, a C++ sampler) needs to read those files to run the fused nets
// on real checkpoints, and to write files stock Torch can torch.load (flat parameter / Adam-state tensors).
//
// The format lives in torch7's File.lua / File.c (third-party, absent from /root/reference, un-pinned; the layout
// below is the 2015-2016 "V 1" object format):
//   object   := int32 type, then
//     0 nil | 1 number: double | 2 string: int32 len + bytes | 5 boolean: int32
//     3 table : int32 index; first occurrence: int32 count, then `count` (key object, value object) pairs
//     4 torch : int32 index; first occurrence: string "V 1", string class name, then the class payload:
//               torch.XTensor : int32 nDim, int64 size[nDim], int64 stride[nDim], int64 storageOffset (1-based),
//                               then the storage as an object (or nil)
//               torch.XStorage: int64 size, raw elements
//               anything else (nn.* modules): ONE object, normally the table of the module's fields
//     6 function (legacy): int32 len + dumped bytes, then the upvalues as an object
//     7/8 recursive function: int32 index; first occurrence: int32 len + bytes, then the upvalues object
//   Repeated indices are references to the first occurrence.  All integers little-endian, long = 8 bytes.
// PARITY UNPINNED: no Torch7 in this image to produce or consume a file; tests build files by the rules above.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "fg_internal.h"

namespace {
enum Kind { K_NIL = 0, K_NUMBER = 1, K_STRING = 2, K_TABLE = 3, K_OBJECT = 4, K_BOOL = 5, K_FUNCTION = 6, K_TENSOR = 16, K_STORAGE = 17 };

struct Obj;
using ObjP = std::shared_ptr<Obj>;
struct Obj {
  int kind = K_NIL;
  double num = 0;
  bool b = false;
  std::string str;                          // string value / class name
  std::vector<std::pair<ObjP, ObjP>> items;  // table
  ObjP payload;                             // generic torch object: its field table; tensor: its storage
  // tensor
  std::vector<int64_t> size, stride;
  int64_t offset = 0;  // 0-based
  // storage
  int elem = 0;  // bytes per element
  char etype = 'f';  // 'f' float, 'd' double, 'l' int64, 'i' int32, 's' int16, 'b' int8/uint8
  std::vector<uint8_t> data;
};

struct Reader {
  FILE* f = nullptr;
  int64_t file_bytes = 0;  // nothing inside the file can be larger than the file: bounds every allocation
  std::map<int, ObjP> memo;
  std::string err;
  bool ok = true;
  bool rd(void* p, size_t n) {
    if (!ok) return false;
    if (fread(p, 1, n, f) != n) {
      ok = false;
      err = "unexpected end of file";
    }
    return ok;
  }
  bool fits(int64_t bytes) {
    if (bytes < 0 || bytes > file_bytes) {
      if (ok) {
        ok = false;
        err = "a length field exceeds the file size (corrupt or not a binary torch.save file)";
      }
      return false;
    }
    return true;
  }
  int32_t i32() {
    int32_t v = 0;
    rd(&v, 4);
    return v;
  }
  int64_t i64() {
    int64_t v = 0;
    rd(&v, 8);
    return v;
  }
  std::string str() {
    const int32_t n = i32();
    if (!ok || n < 0 || n > (1 << 28) || !fits(n)) {
      if (ok) { ok = false; err = "bad string length"; }
      return "";
    }
    std::string s((size_t)n, '\0');
    if (n) rd(&s[0], (size_t)n);
    return s;
  }
  static bool storage_type(const std::string& cls, int* elem, char* et) {
    struct T { const char* n; int e; char t; };
    static const T tab[] = {{"Float", 4, 'f'}, {"Cuda", 4, 'f'}, {"Double", 8, 'd'}, {"CudaDouble", 8, 'd'},
                            {"Long", 8, 'l'}, {"CudaLong", 8, 'l'}, {"Int", 4, 'i'}, {"CudaInt", 4, 'i'},
                            {"Short", 2, 's'}, {"CudaShort", 2, 's'}, {"Char", 1, 'b'}, {"Byte", 1, 'b'},
                            {"CudaChar", 1, 'b'}, {"CudaByte", 1, 'b'}};
    for (const T& t : tab) {
      const std::string a = std::string("torch.") + t.n + "Tensor", s = std::string("torch.") + t.n + "Storage";
      if (cls == a || cls == s) {
        *elem = t.e;
        *et = t.t;
        return true;
      }
    }
    return false;
  }
  ObjP object(int depth = 0) {
    ObjP o = std::make_shared<Obj>();
    if (depth > 512) {
      ok = false;
      err = "nesting too deep";
      return o;
    }
    const int32_t type = i32();
    if (!ok) return o;
    switch (type) {
      case 0: o->kind = K_NIL; return o;
      case 1: o->kind = K_NUMBER; rd(&o->num, 8); return o;
      case 2: o->kind = K_STRING; o->str = str(); return o;
      case 5: o->kind = K_BOOL; o->b = i32() == 1; return o;
      case 6: {  // legacy function: no index
        o->kind = K_FUNCTION;
        o->str = str();
        o->payload = object(depth + 1);
        return o;
      }
      case 3: case 4: case 7: case 8: {
        const int32_t index = i32();
        auto it = memo.find(index);
        if (it != memo.end()) return it->second;
        memo[index] = o;
        if (type == 7 || type == 8) {
          o->kind = K_FUNCTION;
          o->str = str();
          o->payload = object(depth + 1);
        } else if (type == 3) {
          o->kind = K_TABLE;
          const int32_t n = i32();
          if (n < 0 || n > (1 << 26) || !fits((int64_t)n * 8)) {  // every pair needs at least two type words
            if (ok) { ok = false; err = "bad table size"; }
            return o;
          }
          for (int32_t i = 0; i < n && ok; ++i) {
            ObjP k = object(depth + 1);
            ObjP v = object(depth + 1);
            o->items.emplace_back(k, v);
          }
        } else {
          std::string version = str(), cls;
          if (version.compare(0, 2, "V ") == 0) cls = str(); else cls = version;  // pre-versioning files
          o->str = cls;
          int elem;
          char et;
          if (storage_type(cls, &elem, &et) && cls.size() > 6 && cls.compare(cls.size() - 6, 6, "Tensor") == 0) {
            o->kind = K_TENSOR;
            const int32_t nd = i32();
            if (nd < 0 || nd > 16) {
              ok = false;
              err = "bad tensor rank";
              return o;
            }
            o->size.resize(nd);
            o->stride.resize(nd);
            for (int i = 0; i < nd; ++i) o->size[i] = i64();
            for (int i = 0; i < nd; ++i) o->stride[i] = i64();
            o->offset = i64() - 1;
            o->payload = object(depth + 1);
            o->elem = elem;
            o->etype = et;
          } else if (storage_type(cls, &elem, &et)) {
            o->kind = K_STORAGE;
            o->elem = elem;
            o->etype = et;
            const int64_t n = i64();
            if (n < 0 || n > ((int64_t)1 << 36) || !fits(n * elem)) {
              if (ok) { ok = false; err = "bad storage size"; }
              return o;
            }
            o->data.resize((size_t)n * elem);
            if (n) rd(o->data.data(), o->data.size());
          } else {
            o->kind = K_OBJECT;
            o->payload = object(depth + 1);
          }
        }
        return o;
      }
      default:
        ok = false;
        err = "unknown object type " + std::to_string(type);
        return o;
    }
  }
};

const Obj* table_of(const Obj* o) {
  if (!o) return nullptr;
  if (o->kind == K_TABLE) return o;
  if (o->kind == K_OBJECT && o->payload && o->payload->kind == K_TABLE) return o->payload.get();
  return nullptr;
}
const Obj* field(const Obj* o, const std::string& key) {
  const Obj* t = table_of(o);
  if (!t) return nullptr;
  bool numeric = !key.empty();
  for (char ch : key) numeric = numeric && ch >= '0' && ch <= '9';
  for (const auto& kv : t->items) {
    if (kv.first->kind == K_STRING && kv.first->str == key) return kv.second.get();
    if (numeric && kv.first->kind == K_NUMBER && kv.first->num == (double)atoll(key.c_str())) return kv.second.get();
  }
  return nullptr;
}
const Obj* lookup(const Obj* root, const char* path) {
  const Obj* cur = root;
  if (!path || !*path) return cur;
  std::string seg;
  for (const char* p = path;; ++p) {
    if (*p == '.' || *p == '\0') {
      cur = field(cur, seg);
      if (!cur) return nullptr;
      seg.clear();
      if (*p == '\0') break;
    } else {
      seg.push_back(*p);
    }
  }
  return cur;
}
double elem_at(const Obj* st, int64_t i) {
  const uint8_t* p = st->data.data() + (size_t)i * st->elem;
  switch (st->etype) {
    case 'f': { float v; memcpy(&v, p, 4); return v; }
    case 'd': { double v; memcpy(&v, p, 8); return v; }
    case 'l': { int64_t v; memcpy(&v, p, 8); return (double)v; }
    case 'i': { int32_t v; memcpy(&v, p, 4); return (double)v; }
    case 's': { int16_t v; memcpy(&v, p, 2); return (double)v; }
    default: return (double)*p;
  }
}
// element count, saturated so that hostile size fields cannot overflow or drive a multi-hour loop
constexpr int64_t kMaxNumel = (int64_t)1 << 33;
int64_t numel(const Obj* t) {
  if (t->size.empty()) return 0;
  int64_t n = 1;
  for (int64_t s : t->size) {
    if (s < 0) return kMaxNumel + 1;
    if (s == 0) return 0;
    if (n > kMaxNumel / s) return kMaxNumel + 1;
    n *= s;
  }
  return n;
}
// appends the tensor's elements in logical (row-major) order as floats; false if the storage is too small
bool flatten(const Obj* t, std::vector<float>* out) {
  const int64_t n = numel(t);
  if (n == 0) return true;
  if (n > kMaxNumel) return false;
  const Obj* st = t->payload.get();
  if (!st || st->kind != K_STORAGE) return false;
  const int64_t cap = (int64_t)(st->data.size() / st->elem);
  // a tensor of a checkpoint never has more logical elements than its storage holds; an "expanded" (stride-0) view
  // over a tiny storage would otherwise turn a 200-byte file into a multi-GiB allocation
  if (n > cap) return false;
  const int nd = (int)t->size.size();
  std::vector<int64_t> idx(nd, 0);
  for (int64_t i = 0; i < n; ++i) {
    int64_t off = t->offset;
    for (int d = 0; d < nd; ++d) off += idx[d] * t->stride[d];
    if (off < 0 || off >= cap) return false;
    out->push_back((float)elem_at(st, off));
    for (int d = nd - 1; d >= 0; --d) {
      if (++idx[d] < t->size[d]) break;
      idx[d] = 0;
    }
  }
  return true;
}
// nn.Module:parameters() order: containers recurse over self.modules[1..n]; a leaf contributes weight then bias
// The object memo lets a hostile file make a `modules` table contain its own parent (a cycle) or the same subtree
// many times (exponential fan-out): the walk refuses cycles and stops after kMaxModules visits.
constexpr int kMaxModules = 1 << 16;
struct Walk {
  std::vector<const Obj*> path;  // containers on the current recursion path
  int visited = 0;
  bool ok = true;
  bool enter(const Obj* m) {
    if (++visited > kMaxModules || path.size() > 64) return ok = false;
    for (const Obj* p : path)
      if (p == m) return ok = false;
    path.push_back(m);
    return true;
  }
  void leave() { path.pop_back(); }
};
void walk_modules(const Obj* m, std::vector<const Obj*>* leaves, Walk* w) {
  if (!m || !w->ok) return;
  const Obj* mods = field(m, "modules");
  const Obj* mt = table_of(mods);
  if (mt) {
    if (!w->enter(m)) return;
    for (int i = 1; w->ok; ++i) {
      const Obj* child = field(mods, std::to_string(i));
      if (!child) break;
      walk_modules(child, leaves, w);
    }
    w->leave();
    return;
  }
  if (++w->visited > kMaxModules) {
    w->ok = false;
    return;
  }
  leaves->push_back(m);
}
}  // namespace

struct fg_t7 {
  ObjP root;
};
struct fg_t7_writer {
  FILE* f = nullptr;
  struct Entry {
    std::string key;
    int kind;  // K_TENSOR, K_NUMBER, K_STRING
    std::vector<float> data;
    std::vector<int64_t> dims;
    double num = 0;
    std::string str;
  };
  std::vector<Entry> entries;
};

extern "C" {

int fg_t7_open(const char* path, fg_t7** out) {
  if (!path || !out) {
    fg_set_error("fg_t7_open: null argument");
    return FG_ERR_INVALID;
  }
  *out = nullptr;
  Reader r;
  r.f = fopen(path, "rb");
  if (!r.f) {
    fg_set_error("fg_t7_open: cannot open %s", path);
    return FG_ERR_INVALID;
  }
  if (fseek(r.f, 0, SEEK_END) == 0) r.file_bytes = (int64_t)ftell(r.f);
  rewind(r.f);
  ObjP root;
  try {  // nothing may throw across the C ABI (std::bad_alloc on a hostile file, ...)
    root = r.object();
  } catch (const std::exception& e) {
    r.ok = false;
    r.err = std::string("exception while parsing: ") + e.what();
  }
  fclose(r.f);
  if (!r.ok) {
    fg_set_error("fg_t7_open(%s): %s (only the binary torch.save format is supported)", path, r.err.c_str());
    return FG_ERR_INVALID;
  }
  *out = new fg_t7{root};
  return FG_OK;
}
int fg_t7_close(fg_t7* f) {
  delete f;
  return FG_OK;
}
int fg_t7_kind(fg_t7* f, const char* path) {
  const Obj* o = f ? lookup(f->root.get(), path) : nullptr;
  return o ? o->kind : -1;
}
int fg_t7_number(fg_t7* f, const char* path, double* out) {
  const Obj* o = f ? lookup(f->root.get(), path) : nullptr;
  if (!o || !out || (o->kind != K_NUMBER && o->kind != K_BOOL)) {
    fg_set_error("fg_t7_number: %s is not a number", path ? path : "(null)");
    return FG_ERR_INVALID;
  }
  *out = o->kind == K_NUMBER ? o->num : (o->b ? 1.0 : 0.0);
  return FG_OK;
}
int64_t fg_t7_string(fg_t7* f, const char* path, char* dst, int64_t cap) {
  const Obj* o = f ? lookup(f->root.get(), path) : nullptr;
  if (!o || (o->kind != K_STRING && o->kind != K_OBJECT && o->kind != K_TENSOR)) return -1;
  const std::string& s = o->str;  // string value, or the torch class name of an object / tensor
  if (dst && cap > 0) {
    const size_t n = std::min<size_t>(s.size(), (size_t)cap - 1);
    memcpy(dst, s.data(), n);
    dst[n] = '\0';
  }
  return (int64_t)s.size();
}
int64_t fg_t7_tensor(fg_t7* f, const char* path, float* dst, int64_t cap, int64_t* dims8) try {
  const Obj* o = f ? lookup(f->root.get(), path) : nullptr;
  if (!o || o->kind != K_TENSOR) {
    fg_set_error("fg_t7_tensor: %s is not a tensor", path ? path : "(null)");
    return -1;
  }
  const int64_t n = numel(o);
  if (n > kMaxNumel) {
    fg_set_error("fg_t7_tensor: %s has an implausible size (corrupt file?)", path);
    return -1;
  }
  if (dims8) {
    for (int i = 0; i < 8; ++i) dims8[i] = i < (int)o->size.size() ? o->size[i] : 0;
  }
  if (dst) {
    if (cap < n) {
      fg_set_error("fg_t7_tensor: %s has %lld elements, buffer holds %lld", path, (long long)n, (long long)cap);
      return -1;
    }
    std::vector<float> v;
    v.reserve((size_t)n);
    if (!flatten(o, &v)) {
      fg_set_error("fg_t7_tensor: %s references elements outside its storage", path);
      return -1;
    }
    memcpy(dst, v.data(), sizeof(float) * (size_t)n);
  }
  return n;
} catch (const std::exception& e) {  // nothing may throw across the C ABI
  fg_set_error("fg_t7_tensor: %s", e.what());
  return -1;
}
// Flat parameter vector of the nn module tree at `path` in getParameters() order (module order, weight then bias;
// train.lua:151-152).  dst == NULL only counts.  Returns the element count or -1.
int64_t fg_t7_net_params(fg_t7* f, const char* path, float* dst, int64_t cap) try {
  const Obj* o = f ? lookup(f->root.get(), path) : nullptr;
  if (!o || !table_of(o)) {
    fg_set_error("fg_t7_net_params: %s is not a module", path ? path : "(null)");
    return -1;
  }
  std::vector<const Obj*> leaves;
  Walk w;
  walk_modules(o, &leaves, &w);
  if (!w.ok) {
    fg_set_error("fg_t7_net_params: %s is not a module tree (cyclic or implausibly large `modules` tables)", path);
    return -1;
  }
  std::vector<float> flat;
  for (const Obj* m : leaves)
    for (const char* name : {"weight", "bias"}) {
      const Obj* t = field(m, name);
      if (t && t->kind == K_TENSOR && !flatten(t, &flat)) {
        fg_set_error("fg_t7_net_params: a %s tensor references elements outside its storage", name);
        return -1;
      }
    }
  if (dst) {
    if (cap < (int64_t)flat.size()) {
      fg_set_error("fg_t7_net_params: %s has %zu parameters, buffer holds %lld", path, flat.size(), (long long)cap);
      return -1;
    }
    memcpy(dst, flat.data(), sizeof(float) * flat.size());
  }
  return (int64_t)flat.size();
} catch (const std::exception& e) {  // nothing may throw across the C ABI
  fg_set_error("fg_t7_net_params: %s", e.what());
  return -1;
}
// BatchNorm running statistics of the module tree, per BN layer in module order: running_mean[C] then running_var[C]
// (2015 `nn` stored running_std = 1/sqrt(var + eps) instead; it is converted back using the module's eps).
int64_t fg_t7_net_bn_state(fg_t7* f, const char* path, float* dst, int64_t cap) try {
  const Obj* o = f ? lookup(f->root.get(), path) : nullptr;
  if (!o || !table_of(o)) {
    fg_set_error("fg_t7_net_bn_state: %s is not a module", path ? path : "(null)");
    return -1;
  }
  std::vector<const Obj*> leaves;
  Walk w;
  walk_modules(o, &leaves, &w);
  if (!w.ok) {
    fg_set_error("fg_t7_net_bn_state: %s is not a module tree (cyclic or implausibly large `modules` tables)", path);
    return -1;
  }
  std::vector<float> flat;
  for (const Obj* m : leaves) {
    const Obj* rm = field(m, "running_mean");
    if (!rm || rm->kind != K_TENSOR) continue;
    if (!flatten(rm, &flat)) return -1;
    const Obj* rv = field(m, "running_var");
    if (rv && rv->kind == K_TENSOR) {
      if (!flatten(rv, &flat)) return -1;
    } else {
      const Obj* rs = field(m, "running_std");
      if (!rs || rs->kind != K_TENSOR) {
        fg_set_error("fg_t7_net_bn_state: BatchNorm module without running_var / running_std");
        return -1;
      }
      const Obj* e = field(m, "eps");
      const double eps = e && e->kind == K_NUMBER ? e->num : 1e-5;
      const size_t at = flat.size();
      if (!flatten(rs, &flat)) return -1;
      for (size_t i = at; i < flat.size(); ++i) flat[i] = (float)(1.0 / ((double)flat[i] * flat[i]) - eps);
    }
  }
  if (dst) {
    if (cap < (int64_t)flat.size()) {
      fg_set_error("fg_t7_net_bn_state: buffer too small");
      return -1;
    }
    memcpy(dst, flat.data(), sizeof(float) * flat.size());
  }
  return (int64_t)flat.size();
} catch (const std::exception& e) {  // nothing may throw across the C ABI
  fg_set_error("fg_t7_net_bn_state: %s", e.what());
  return -1;
}
// "nn.Sequential{nn.Copy,nn.Sequential{nn.Linear,...},nn.Copy}" -- lets a host check it loads the architecture it expects
int64_t fg_t7_net_describe(fg_t7* f, const char* path, char* dst, int64_t cap) try {
  const Obj* o = f ? lookup(f->root.get(), path) : nullptr;
  if (!o) return -1;
  std::string s;
  struct Rec {
    static void go(const Obj* m, std::string* s, Walk* w) {
      if (!m || !w->ok) return;
      *s += m->kind == K_OBJECT ? m->str : "?";
      const Obj* mods = field(m, "modules");
      if (table_of(mods)) {
        if (!w->enter(m)) return;
        *s += "{";
        for (int i = 1; w->ok; ++i) {
          const Obj* ch = field(mods, std::to_string(i));
          if (!ch) break;
          if (i > 1) *s += ",";
          go(ch, s, w);
        }
        *s += "}";
        w->leave();
      } else if (++w->visited > kMaxModules) {
        w->ok = false;
      }
    }
  };
  Walk w;
  Rec::go(o, &s, &w);
  if (!w.ok) {
    fg_set_error("fg_t7_net_describe: %s is not a module tree (cyclic or implausibly large `modules` tables)", path);
    return -1;
  }
  if (dst && cap > 0) {
    const size_t n = std::min<size_t>(s.size(), (size_t)cap - 1);
    memcpy(dst, s.data(), n);
    dst[n] = '\0';
  }
  return (int64_t)s.size();
} catch (const std::exception& e) {  // nothing may throw across the C ABI
  fg_set_error("fg_t7_net_describe: %s", e.what());
  return -1;
}

// ---- writer: one root table {key = FloatTensor | number | string}, loadable with stock torch.load ------------
int fg_t7_writer_open(const char* path, fg_t7_writer** out) {
  if (!path || !out) {
    fg_set_error("fg_t7_writer_open: null argument");
    return FG_ERR_INVALID;
  }
  *out = nullptr;
  FILE* f = fopen(path, "wb");
  if (!f) {
    fg_set_error("fg_t7_writer_open: cannot create %s", path);
    return FG_ERR_INVALID;
  }
  fg_t7_writer* w = new fg_t7_writer();
  w->f = f;
  *out = w;
  return FG_OK;
}
int fg_t7_writer_add_tensor(fg_t7_writer* w, const char* key, const float* data, const int64_t* dims, int ndim) try {
  if (!w || !key || !data || !dims || ndim < 1 || ndim > 8) {
    fg_set_error("fg_t7_writer_add_tensor: bad arguments");
    return FG_ERR_INVALID;
  }
  fg_t7_writer::Entry e;
  e.key = key;
  e.kind = K_TENSOR;
  int64_t n = 1;
  for (int i = 0; i < ndim; ++i) {
    if (dims[i] < 1) {
      fg_set_error("fg_t7_writer_add_tensor: non-positive dimension");
      return FG_ERR_INVALID;
    }
    e.dims.push_back(dims[i]);
    n *= dims[i];
  }
  e.data.assign(data, data + n);
  w->entries.push_back(std::move(e));
  return FG_OK;
} catch (const std::exception& e) {  // nothing may throw across the C ABI
  fg_set_error("fg_t7_writer_add_tensor: %s", e.what());
  return FG_ERR_INVALID;
}
int fg_t7_writer_add_number(fg_t7_writer* w, const char* key, double v) {
  if (!w || !key) return FG_ERR_INVALID;
  fg_t7_writer::Entry e;
  e.key = key;
  e.kind = K_NUMBER;
  e.num = v;
  w->entries.push_back(std::move(e));
  return FG_OK;
}
int fg_t7_writer_add_string(fg_t7_writer* w, const char* key, const char* s) {
  if (!w || !key || !s) return FG_ERR_INVALID;
  fg_t7_writer::Entry e;
  e.key = key;
  e.kind = K_STRING;
  e.str = s;
  w->entries.push_back(std::move(e));
  return FG_OK;
}
int fg_t7_writer_close(fg_t7_writer* w) {
  if (!w) return FG_OK;
  FILE* f = w->f;
  bool ok = true;
  auto wi32 = [&](int32_t v) { ok = ok && fwrite(&v, 4, 1, f) == 1; };
  auto wi64 = [&](int64_t v) { ok = ok && fwrite(&v, 8, 1, f) == 1; };
  auto wstr = [&](const std::string& s) {
    wi32((int32_t)s.size());
    if (!s.empty()) ok = ok && fwrite(s.data(), 1, s.size(), f) == s.size();
  };
  int32_t next_index = 1;
  wi32(3);  // root table
  wi32(next_index++);
  wi32((int32_t)w->entries.size());
  for (const auto& e : w->entries) {
    wi32(2);
    wstr(e.key);
    if (e.kind == K_NUMBER) {
      wi32(1);
      ok = ok && fwrite(&e.num, 8, 1, f) == 1;
    } else if (e.kind == K_STRING) {
      wi32(2);
      wstr(e.str);
    } else {
      wi32(4);
      wi32(next_index++);
      wstr("V 1");
      wstr("torch.FloatTensor");
      wi32((int32_t)e.dims.size());
      for (int64_t d : e.dims) wi64(d);
      int64_t stride = 1;
      std::vector<int64_t> st(e.dims.size());
      for (int i = (int)e.dims.size() - 1; i >= 0; --i) {
        st[i] = stride;
        stride *= e.dims[i];
      }
      for (int64_t s : st) wi64(s);
      wi64(1);  // storageOffset, 1-based
      wi32(4);
      wi32(next_index++);
      wstr("V 1");
      wstr("torch.FloatStorage");
      wi64((int64_t)e.data.size());
      ok = ok && fwrite(e.data.data(), sizeof(float), e.data.size(), f) == e.data.size();
    }
  }
  ok = (fclose(f) == 0) && ok;
  delete w;
  if (!ok) {
    fg_set_error("fg_t7_writer_close: write failed");
    return FG_ERR_INVALID;
  }
  return FG_OK;
}

}  // extern "C"
