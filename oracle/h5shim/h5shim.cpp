// h5shim.cpp -- TEST / BASELINE INFRASTRUCTURE (see hdf5.h in this directory).
//
// The HDF5 entry points the reference calls (src/hdf5_interface.cpp, src/state_point.cpp,
// include/openmc/bank_io.h, src/track_output.cpp, src/collision_track.cpp, src/plot.cpp),
// implemented over an in-memory tree that is loaded from / stored to a private container
// file ("OMH5" format, written and read by oracle/omh5.py as well).  Nothing here is
// HDF5 code and nothing of the product links it.
//
// Container layout (little endian):
//   file  := "OMH5\1\0\0\0" node
//   node  := u8 kind(0 group,1 dataset) str name u32 n_attr {str name data}*
//            ( group: u32 n_child node* | dataset: data )
//   data  := type u8 scalar u32 rank u64 dims[rank] u64 n_bytes bytes
//   type  := u8 class(0 int,1 uint,2 float,3 string,4 compound) u64 size
//            ( compound: u32 n_member {str name u64 offset type}* )
//   str   := u32 length bytes
#include "hdf5.h"
#include "hdf5_hl.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace {

struct Type {
  enum Cls : uint8_t { INT = 0, UINT = 1, FLOAT = 2, STRING = 3, COMPOUND = 4 };
  struct Member;
  Cls cls = INT;
  size_t size = 0;
  std::vector<Member> members;
};
struct Type::Member {
  std::string name;
  size_t offset;
  Type type;
};

struct Data {
  Type type;
  bool scalar = true;
  std::vector<hsize_t> dims;
  std::vector<uint8_t> bytes;
  size_t count() const
  {
    size_t n = 1;
    for (auto d : dims) n *= d;
    return n;
  }
};

struct Node {
  bool is_dataset = false;
  std::string name;
  Node* parent = nullptr;
  std::map<std::string, Data> attrs;
  std::map<std::string, std::unique_ptr<Node>> children;   // name order = H5_INDEX_NAME
  Data data;
};

struct File {
  std::string path;
  bool writable = false;
  bool dirty = false;
  int refs = 0;
  std::unique_ptr<Node> root;
};

struct Space {
  bool scalar = true;
  std::vector<hsize_t> dims;
  bool selected = false;
  std::vector<hsize_t> start, count;
  size_t n_selected() const
  {
    size_t n = 1;
    if (selected) for (auto c : count) n *= c;
    else for (auto d : dims) n *= d;
    return n;
  }
};

struct Handle {
  H5I_type_t kind = H5I_BADID;
  std::shared_ptr<File> file;   // FILE, GROUP, DATASET, ATTR
  Node* node = nullptr;         // GROUP, DATASET, ATTR owner
  std::string attr;             // ATTR
  Space space;                  // DATASPACE
  Type type;                    // DATATYPE
  hid_t plist_class = 0;        // GENPROP_LST
};

std::mutex g_mutex;
std::map<hid_t, Handle> g_handles;
hid_t g_next = 1000;

Type predefined(hid_t id)
{
  Type t;
  switch (id) {
  case 1: t.cls = Type::INT; t.size = 1; break;
  case 2: t.cls = Type::UINT; t.size = 1; break;
  case 3: t.cls = Type::INT; t.size = 2; break;
  case 4: t.cls = Type::UINT; t.size = 2; break;
  case 5: t.cls = Type::INT; t.size = 4; break;
  case 6: t.cls = Type::UINT; t.size = 4; break;
  case 7: t.cls = Type::INT; t.size = 8; break;
  case 8: t.cls = Type::UINT; t.size = 8; break;
  case 9: t.cls = Type::FLOAT; t.size = 4; break;
  case 10: t.cls = Type::FLOAT; t.size = 8; break;
  case 11: t.cls = Type::STRING; t.size = 1; break;
  default: t.size = 0;
  }
  return t;
}

bool get_type(hid_t id, Type& out)
{
  if (id >= 1 && id <= 11) { out = predefined(id); return true; }
  auto it = g_handles.find(id);
  if (it == g_handles.end() || it->second.kind != H5I_DATATYPE) return false;
  out = it->second.type;
  return true;
}

hid_t new_handle(Handle&& h)
{
  hid_t id = g_next++;
  g_handles[id] = std::move(h);
  return id;
}

Handle* find(hid_t id)
{
  auto it = g_handles.find(id);
  return it == g_handles.end() ? nullptr : &it->second;
}

// ---- path walking --------------------------------------------------------------------------
Node* walk(Handle* loc, const char* path, bool create_leaf_group = false, bool* created = nullptr)
{
  if (!loc || !loc->node) return nullptr;
  Node* cur = loc->node;
  std::string p(path ? path : "");
  size_t i = 0;
  if (!p.empty() && p[0] == '/') { cur = loc->file->root.get(); i = 1; }
  while (i <= p.size()) {
    size_t j = p.find('/', i);
    if (j == std::string::npos) j = p.size();
    std::string part = p.substr(i, j - i);
    i = j + 1;
    if (part.empty() || part == ".") { if (j >= p.size()) break; continue; }
    auto it = cur->children.find(part);
    if (it == cur->children.end()) {
      if (create_leaf_group && j >= p.size()) {
        auto n = std::make_unique<Node>();
        n->name = part;
        n->parent = cur;
        Node* raw = n.get();
        cur->children[part] = std::move(n);
        if (created) *created = true;
        return raw;
      }
      return nullptr;
    }
    cur = it->second.get();
    if (j >= p.size()) break;
  }
  return cur;
}

std::string full_name(Node* n)
{
  if (!n->parent) return "/";
  std::string s;
  for (Node* c = n; c->parent; c = c->parent) s = "/" + c->name + s;
  return s;
}

// ---- serialisation -------------------------------------------------------------------------
struct Writer {
  FILE* f;
  void raw(const void* p, size_t n) { if (n) fwrite(p, 1, n, f); }
  void u8(uint8_t v) { raw(&v, 1); }
  void u32(uint32_t v) { raw(&v, 4); }
  void u64(uint64_t v) { raw(&v, 8); }
  void str(const std::string& s) { u32((uint32_t)s.size()); raw(s.data(), s.size()); }
  void type(const Type& t)
  {
    u8(t.cls); u64(t.size);
    if (t.cls == Type::COMPOUND) {
      u32((uint32_t)t.members.size());
      for (auto& m : t.members) { str(m.name); u64(m.offset); type(m.type); }
    }
  }
  void data(const Data& d)
  {
    type(d.type); u8(d.scalar ? 1 : 0); u32((uint32_t)d.dims.size());
    for (auto x : d.dims) u64(x);
    u64(d.bytes.size()); raw(d.bytes.data(), d.bytes.size());
  }
  void node(const Node& n)
  {
    u8(n.is_dataset ? 1 : 0); str(n.name); u32((uint32_t)n.attrs.size());
    for (auto& a : n.attrs) { str(a.first); data(a.second); }
    if (n.is_dataset) data(n.data);
    else {
      u32((uint32_t)n.children.size());
      for (auto& c : n.children) node(*c.second);
    }
  }
};

struct Reader {
  FILE* f;
  bool ok = true;
  void raw(void* p, size_t n) { if (n && fread(p, 1, n, f) != n) ok = false; }
  uint8_t u8() { uint8_t v = 0; raw(&v, 1); return v; }
  uint32_t u32() { uint32_t v = 0; raw(&v, 4); return v; }
  uint64_t u64() { uint64_t v = 0; raw(&v, 8); return v; }
  std::string str() { uint32_t n = u32(); std::string s(ok ? n : 0, '\0'); raw(&s[0], s.size()); return s; }
  Type type()
  {
    Type t; t.cls = (Type::Cls)u8(); t.size = u64();
    if (t.cls == Type::COMPOUND) {
      uint32_t n = u32();
      for (uint32_t i = 0; ok && i < n; ++i) {
        Type::Member m; m.name = str(); m.offset = u64(); m.type = type();
        t.members.push_back(std::move(m));
      }
    }
    return t;
  }
  Data data()
  {
    Data d; d.type = type(); d.scalar = u8() != 0;
    uint32_t r = u32();
    for (uint32_t i = 0; ok && i < r; ++i) d.dims.push_back(u64());
    uint64_t nb = u64();
    if (ok) { d.bytes.resize(nb); raw(d.bytes.data(), nb); }
    return d;
  }
  std::unique_ptr<Node> node(Node* parent)
  {
    auto n = std::make_unique<Node>();
    n->parent = parent;
    n->is_dataset = u8() != 0;
    n->name = str();
    uint32_t na = u32();
    for (uint32_t i = 0; ok && i < na; ++i) { std::string k = str(); n->attrs[k] = data(); }
    if (n->is_dataset) n->data = data();
    else {
      uint32_t nc = u32();
      for (uint32_t i = 0; ok && i < nc; ++i) {
        auto c = node(n.get());
        std::string k = c->name;
        n->children[k] = std::move(c);
      }
    }
    return n;
  }
};

const char MAGIC[8] = {'O', 'M', 'H', '5', 1, 0, 0, 0};

bool save(const File& file)
{
  FILE* f = fopen(file.path.c_str(), "wb");
  if (!f) return false;
  fwrite(MAGIC, 1, 8, f);
  Writer w{f};
  w.node(*file.root);
  return fclose(f) == 0;
}

bool load(File& file)
{
  FILE* f = fopen(file.path.c_str(), "rb");
  if (!f) return false;
  char m[8];
  if (fread(m, 1, 8, f) != 8 || memcmp(m, MAGIC, 8) != 0) { fclose(f); return false; }
  Reader r{f};
  file.root = r.node(nullptr);
  fclose(f);
  return r.ok;
}

// ---- element conversion --------------------------------------------------------------------
template<class T> T load_as(const uint8_t* p) { T v; memcpy(&v, p, sizeof v); return v; }

void convert(const Type& st, const uint8_t* s, const Type& dt, uint8_t* d)
{
  if (dt.cls == Type::COMPOUND) {
    if (st.cls != Type::COMPOUND) return;
    for (auto& dm : dt.members)
      for (auto& sm : st.members)
        if (sm.name == dm.name) { convert(sm.type, s + sm.offset, dm.type, d + dm.offset); break; }
    return;
  }
  if (dt.cls == Type::STRING) {
    memset(d, 0, dt.size);
    if (st.cls == Type::STRING) memcpy(d, s, st.size < dt.size ? st.size : dt.size);
    return;
  }
  if (st.cls == dt.cls && st.size == dt.size) { memcpy(d, s, dt.size); return; }
  // numeric conversion through long double / 64-bit integers
  bool s_float = st.cls == Type::FLOAT;
  double fv = 0; int64_t iv = 0; uint64_t uv = 0;
  if (s_float) fv = st.size == 4 ? (double)load_as<float>(s) : load_as<double>(s);
  else if (st.cls == Type::INT) {
    switch (st.size) { case 1: iv = load_as<int8_t>(s); break; case 2: iv = load_as<int16_t>(s); break;
      case 4: iv = load_as<int32_t>(s); break; default: iv = load_as<int64_t>(s); }
    uv = (uint64_t)iv; fv = (double)iv;
  } else if (st.cls == Type::UINT) {
    switch (st.size) { case 1: uv = load_as<uint8_t>(s); break; case 2: uv = load_as<uint16_t>(s); break;
      case 4: uv = load_as<uint32_t>(s); break; default: uv = load_as<uint64_t>(s); }
    iv = (int64_t)uv; fv = (double)uv;
  } else return;
  if (dt.cls == Type::FLOAT) {
    if (dt.size == 4) { float v = (float)fv; memcpy(d, &v, 4); } else memcpy(d, &fv, 8);
  } else {
    if (s_float) { iv = (int64_t)fv; uv = (uint64_t)iv; }
    uint64_t bits = dt.cls == Type::INT ? (uint64_t)iv : uv;
    memcpy(d, &bits, dt.size);   // little endian: low bytes first
  }
}

// linear element offsets of a selection, in row-major order
void selection_offsets(const Space& sp, size_t n_total_default, std::vector<size_t>& out)
{
  if (!sp.selected) {
    size_t n = sp.scalar && sp.dims.empty() ? 1 : sp.n_selected();
    if (n_total_default && sp.dims.empty() && !sp.scalar) n = n_total_default;
    out.resize(n);
    for (size_t i = 0; i < n; ++i) out[i] = i;
    return;
  }
  size_t rank = sp.dims.size();
  std::vector<size_t> stride(rank, 1);
  for (size_t k = rank; k-- > 1;) stride[k - 1] = stride[k] * sp.dims[k];
  size_t n = sp.n_selected();
  out.resize(n);
  std::vector<hsize_t> idx(rank, 0);
  for (size_t i = 0; i < n; ++i) {
    size_t off = 0;
    for (size_t k = 0; k < rank; ++k) off += (sp.start[k] + idx[k]) * stride[k];
    out[i] = off;
    for (size_t k = rank; k-- > 0;) {
      if (++idx[k] < sp.count[k]) break;
      idx[k] = 0;
    }
  }
}

Space space_of(const Data& d)
{
  Space s; s.scalar = d.scalar; s.dims = d.dims; return s;
}

bool get_space(hid_t id, const Data& d, Space& out)
{
  if (id == H5S_ALL) { out = space_of(d); return true; }
  Handle* h = find(id);
  if (!h || h->kind != H5I_DATASPACE) return false;
  out = h->space;
  return true;
}

herr_t transfer(Data& d, hid_t mem_type, hid_t mem_space, hid_t file_space, void* rbuf, const void* wbuf)
{
  Type mt;
  if (!get_type(mem_type, mt)) return -1;
  Space ms, fs;
  if (!get_space(file_space, d, fs)) return -1;
  if (mem_space == H5S_ALL) {
    // memory = a dense block of as many elements as the file selection holds
    ms.scalar = false; ms.dims = {(hsize_t)fs.n_selected()};
  } else if (!get_space(mem_space, d, ms)) return -1;
  std::vector<size_t> mo, fo;
  selection_offsets(ms, 0, mo);
  selection_offsets(fs, 0, fo);
  size_t n = mo.size() < fo.size() ? mo.size() : fo.size();
  size_t fsz = d.type.size, msz = mt.size;
  if (rbuf) {
    uint8_t* out = static_cast<uint8_t*>(rbuf);
    for (size_t i = 0; i < n; ++i) {
      if ((fo[i] + 1) * fsz > d.bytes.size()) return -1;
      convert(d.type, d.bytes.data() + fo[i] * fsz, mt, out + mo[i] * msz);
    }
  } else {
    const uint8_t* in = static_cast<const uint8_t*>(wbuf);
    for (size_t i = 0; i < n; ++i) {
      if ((fo[i] + 1) * fsz > d.bytes.size()) return -1;
      convert(mt, in + mo[i] * msz, d.type, d.bytes.data() + fo[i] * fsz);
    }
  }
  return 0;
}

Data make_data(const Type& t, const Space& s)
{
  Data d; d.type = t; d.scalar = s.scalar; d.dims = s.dims;
  d.bytes.assign(d.count() * t.size, 0);
  return d;
}

void release_file(std::shared_ptr<File>& f)
{
  if (!f) return;
  if (--f->refs == 0 && f->writable && f->dirty) { save(*f); f->dirty = false; }
}

hid_t open_node(hid_t loc, const char* name, int want /* -1 any, 0 group, 1 dataset */)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  Handle* h = find(loc);
  Node* n = walk(h, name);
  if (!n) return -1;
  if (want == 0 && n->is_dataset) return -1;
  if (want == 1 && !n->is_dataset) return -1;
  Handle o; o.kind = n->is_dataset ? H5I_DATASET : H5I_GROUP; o.file = h->file; o.node = n;
  o.file->refs++;
  return new_handle(std::move(o));
}

herr_t close_handle(hid_t id)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  auto it = g_handles.find(id);
  if (it == g_handles.end()) return (id >= 0 && id <= 22) ? 0 : -1;
  if (it->second.file) release_file(it->second.file);
  g_handles.erase(it);
  return 0;
}

Node* nth_child(Handle* h, const char* group_name, hsize_t n)
{
  Node* g = walk(h, group_name);
  if (!g || n >= g->children.size()) return nullptr;
  auto it = g->children.begin();
  std::advance(it, (long)n);
  return it->second.get();
}

} // namespace

extern "C" {

// ---- files ---------------------------------------------------------------------------------
hid_t H5Fcreate(const char* name, unsigned flags, hid_t, hid_t)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  if (flags & H5F_ACC_EXCL) {
    FILE* f = fopen(name, "rb");
    if (f) { fclose(f); return -1; }
  }
  auto file = std::make_shared<File>();
  file->path = name; file->writable = true; file->dirty = true; file->refs = 1;
  file->root = std::make_unique<Node>();
  Handle h; h.kind = H5I_FILE; h.file = file; h.node = file->root.get();
  return new_handle(std::move(h));
}

hid_t H5Fopen(const char* name, unsigned flags, hid_t)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  auto file = std::make_shared<File>();
  file->path = name; file->writable = (flags & H5F_ACC_RDWR) != 0; file->refs = 1;
  if (!load(*file)) return -1;
  Handle h; h.kind = H5I_FILE; h.file = file; h.node = file->root.get();
  return new_handle(std::move(h));
}

herr_t H5Fclose(hid_t file) { return close_handle(file); }

hid_t H5Fget_access_plist(hid_t)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  Handle h; h.kind = H5I_GENPROP_LST; h.plist_class = H5P_FILE_ACCESS;
  return new_handle(std::move(h));
}

// ---- groups / objects ----------------------------------------------------------------------
hid_t H5Gcreate(hid_t loc, const char* name, hid_t, hid_t, hid_t)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  Handle* h = find(loc);
  bool created = false;
  Node* n = walk(h, name, true, &created);
  if (!n || !created) return -1;
  h->file->dirty = true;
  Handle o; o.kind = H5I_GROUP; o.file = h->file; o.node = n; o.file->refs++;
  return new_handle(std::move(o));
}
hid_t H5Gopen(hid_t loc, const char* name, hid_t) { return open_node(loc, name, 0); }
herr_t H5Gclose(hid_t g) { return close_handle(g); }
hid_t H5Oopen(hid_t loc, const char* name, hid_t) { return open_node(loc, name, -1); }
herr_t H5Oclose(hid_t o) { return close_handle(o); }

herr_t H5Gget_info(hid_t group, H5G_info_t* info)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  Handle* h = find(group);
  if (!h || !h->node) return -1;
  info->nlinks = h->node->children.size();
  return 0;
}

herr_t H5Oget_info_by_idx(hid_t loc, const char* group_name, H5_index_t, H5_iter_order_t, hsize_t n,
  H5O_info_t* info, hid_t)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  Node* c = nth_child(find(loc), group_name, n);
  if (!c) return -1;
  info->type = c->is_dataset ? H5O_TYPE_DATASET : H5O_TYPE_GROUP;
  return 0;
}

long H5Lget_name_by_idx(hid_t loc, const char* group_name, H5_index_t, H5_iter_order_t, hsize_t n,
  char* name, size_t size, hid_t)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  Node* c = nth_child(find(loc), group_name, n);
  if (!c) return -1;
  if (name && size) {
    size_t k = c->name.size() < size - 1 ? c->name.size() : size - 1;
    memcpy(name, c->name.data(), k);
    name[k] = '\0';
  }
  return (long)c->name.size();
}

htri_t H5LTpath_valid(hid_t loc, const char* path, int)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  Handle* h = find(loc);
  if (!h || !h->node) return -1;
  return walk(h, path) ? 1 : 0;
}

// ---- identifiers ---------------------------------------------------------------------------
long H5Iget_name(hid_t id, char* name, size_t size)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  Handle* h = find(id);
  if (!h || !h->node) return -1;
  std::string s = full_name(h->node);
  if (name && size) {
    size_t k = s.size() < size - 1 ? s.size() : size - 1;
    memcpy(name, s.data(), k);
    name[k] = '\0';
  }
  return (long)s.size();
}

H5I_type_t H5Iget_type(hid_t id)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  if (id >= 1 && id <= 11) return H5I_DATATYPE;
  Handle* h = find(id);
  return h ? h->kind : H5I_BADID;
}

hid_t H5Iget_file_id(hid_t id)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  Handle* h = find(id);
  if (!h || !h->file) return -1;
  Handle o; o.kind = H5I_FILE; o.file = h->file; o.node = h->file->root.get(); o.file->refs++;
  return new_handle(std::move(o));
}

// ---- datasets ------------------------------------------------------------------------------
hid_t H5Dcreate(hid_t loc, const char* name, hid_t type, hid_t space, hid_t, hid_t, hid_t)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  Handle* h = find(loc);
  Handle* sp = find(space);
  Type t;
  if (!h || !sp || sp->kind != H5I_DATASPACE || !get_type(type, t)) return -1;
  bool created = false;
  Node* n = walk(h, name, true, &created);
  if (!n || !created) return -1;
  n->is_dataset = true;
  n->data = make_data(t, sp->space);
  h->file->dirty = true;
  Handle o; o.kind = H5I_DATASET; o.file = h->file; o.node = n; o.file->refs++;
  return new_handle(std::move(o));
}
hid_t H5Dopen(hid_t loc, const char* name, hid_t) { return open_node(loc, name, 1); }
herr_t H5Dclose(hid_t d) { return close_handle(d); }

hid_t H5Dget_space(hid_t dset)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  Handle* h = find(dset);
  if (!h || h->kind != H5I_DATASET) return -1;
  Handle o; o.kind = H5I_DATASPACE; o.space = space_of(h->node->data);
  return new_handle(std::move(o));
}

hid_t H5Dget_type(hid_t dset)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  Handle* h = find(dset);
  if (!h || h->kind != H5I_DATASET) return -1;
  Handle o; o.kind = H5I_DATATYPE; o.type = h->node->data.type;
  return new_handle(std::move(o));
}

herr_t H5Dread(hid_t dset, hid_t mem_type, hid_t mem_space, hid_t file_space, hid_t, void* buf)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  Handle* h = find(dset);
  if (!h || h->kind != H5I_DATASET) return -1;
  return transfer(h->node->data, mem_type, mem_space, file_space, buf, nullptr);
}

herr_t H5Dwrite(hid_t dset, hid_t mem_type, hid_t mem_space, hid_t file_space, hid_t, const void* buf)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  Handle* h = find(dset);
  if (!h || h->kind != H5I_DATASET) return -1;
  h->file->dirty = true;
  return transfer(h->node->data, mem_type, mem_space, file_space, nullptr, buf);
}

// ---- attributes ----------------------------------------------------------------------------
htri_t H5Aexists_by_name(hid_t loc, const char* obj_name, const char* attr_name, hid_t)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  Node* n = walk(find(loc), obj_name);
  if (!n) return -1;
  return n->attrs.count(attr_name) ? 1 : 0;
}

hid_t H5Aopen(hid_t obj, const char* name, hid_t)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  Handle* h = find(obj);
  if (!h || !h->node || !h->node->attrs.count(name)) return -1;
  Handle o; o.kind = H5I_ATTR; o.file = h->file; o.node = h->node; o.attr = name; o.file->refs++;
  return new_handle(std::move(o));
}

hid_t H5Acreate(hid_t obj, const char* name, hid_t type, hid_t space, hid_t, hid_t)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  Handle* h = find(obj);
  Handle* sp = find(space);
  Type t;
  if (!h || !h->node || !sp || sp->kind != H5I_DATASPACE || !get_type(type, t)) return -1;
  h->node->attrs[name] = make_data(t, sp->space);
  h->file->dirty = true;
  Handle o; o.kind = H5I_ATTR; o.file = h->file; o.node = h->node; o.attr = name; o.file->refs++;
  return new_handle(std::move(o));
}
herr_t H5Aclose(hid_t a) { return close_handle(a); }

hid_t H5Aget_space(hid_t attr)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  Handle* h = find(attr);
  if (!h || h->kind != H5I_ATTR) return -1;
  Handle o; o.kind = H5I_DATASPACE; o.space = space_of(h->node->attrs[h->attr]);
  return new_handle(std::move(o));
}

hid_t H5Aget_type(hid_t attr)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  Handle* h = find(attr);
  if (!h || h->kind != H5I_ATTR) return -1;
  Handle o; o.kind = H5I_DATATYPE; o.type = h->node->attrs[h->attr].type;
  return new_handle(std::move(o));
}

herr_t H5Aread(hid_t attr, hid_t mem_type, void* buf)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  Handle* h = find(attr);
  if (!h || h->kind != H5I_ATTR) return -1;
  return transfer(h->node->attrs[h->attr], mem_type, H5S_ALL, H5S_ALL, buf, nullptr);
}

herr_t H5Awrite(hid_t attr, hid_t mem_type, const void* buf)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  Handle* h = find(attr);
  if (!h || h->kind != H5I_ATTR) return -1;
  h->file->dirty = true;
  return transfer(h->node->attrs[h->attr], mem_type, H5S_ALL, H5S_ALL, nullptr, buf);
}

// ---- dataspaces ----------------------------------------------------------------------------
hid_t H5Screate(H5S_class_t cls)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  Handle o; o.kind = H5I_DATASPACE; o.space.scalar = cls == H5S_SCALAR;
  return new_handle(std::move(o));
}

hid_t H5Screate_simple(int rank, const hsize_t* dims, const hsize_t*)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  Handle o; o.kind = H5I_DATASPACE; o.space.scalar = false;
  o.space.dims.assign(dims, dims + rank);
  return new_handle(std::move(o));
}
herr_t H5Sclose(hid_t s) { return close_handle(s); }

int H5Sget_simple_extent_ndims(hid_t space)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  Handle* h = find(space);
  if (!h || h->kind != H5I_DATASPACE) return -1;
  return (int)h->space.dims.size();
}

int H5Sget_simple_extent_dims(hid_t space, hsize_t* dims, hsize_t* maxdims)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  Handle* h = find(space);
  if (!h || h->kind != H5I_DATASPACE) return -1;
  for (size_t i = 0; i < h->space.dims.size(); ++i) {
    if (dims) dims[i] = h->space.dims[i];
    if (maxdims) maxdims[i] = h->space.dims[i];
  }
  return (int)h->space.dims.size();
}

herr_t H5Sselect_hyperslab(hid_t space, H5S_seloper_t, const hsize_t* start, const hsize_t* stride,
  const hsize_t* count, const hsize_t* block)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  Handle* h = find(space);
  if (!h || h->kind != H5I_DATASPACE || stride || block) return -1;   // unit stride / block only
  size_t r = h->space.dims.size();
  h->space.selected = true;
  h->space.start.assign(start, start + r);
  h->space.count.assign(count, count + r);
  return 0;
}

herr_t H5Soffset_simple(hid_t space, const hssize_t* offset)
{
  // the reference (src/plot.cpp:1025) sets an absolute offset of a selection made at start 0
  std::lock_guard<std::mutex> lock(g_mutex);
  Handle* h = find(space);
  if (!h || h->kind != H5I_DATASPACE || !h->space.selected) return -1;
  for (size_t i = 0; i < h->space.start.size(); ++i) h->space.start[i] = (hsize_t)offset[i];
  return 0;
}

// ---- datatypes -----------------------------------------------------------------------------
hid_t H5Tcreate(H5T_class_t cls, size_t size)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  if (cls != H5T_COMPOUND) return -1;
  Handle o; o.kind = H5I_DATATYPE; o.type.cls = Type::COMPOUND; o.type.size = size;
  return new_handle(std::move(o));
}

hid_t H5Tcopy(hid_t type)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  Handle o; o.kind = H5I_DATATYPE;
  if (!get_type(type, o.type)) return -1;
  return new_handle(std::move(o));
}
herr_t H5Tclose(hid_t t) { return close_handle(t); }

herr_t H5Tinsert(hid_t type, const char* name, size_t offset, hid_t member)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  Handle* h = find(type);
  Type m;
  if (!h || h->kind != H5I_DATATYPE || h->type.cls != Type::COMPOUND || !get_type(member, m)) return -1;
  h->type.members.push_back({name, offset, m});
  return 0;
}

herr_t H5Tset_size(hid_t type, size_t size)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  Handle* h = find(type);
  if (!h || h->kind != H5I_DATATYPE) return -1;
  h->type.size = size;
  return 0;
}

size_t H5Tget_size(hid_t type)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  Type t;
  return get_type(type, t) ? t.size : 0;
}

herr_t H5Tset_strpad(hid_t, H5T_str_t) { return 0; }

int H5Tget_nmembers(hid_t type)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  Type t;
  if (!get_type(type, t) || t.cls != Type::COMPOUND) return -1;
  return (int)t.members.size();
}

char* H5Tget_member_name(hid_t type, unsigned idx)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  Type t;
  if (!get_type(type, t) || idx >= t.members.size()) return nullptr;
  return strdup(t.members[idx].name.c_str());
}

herr_t H5free_memory(void* p) { free(p); return 0; }

// ---- property lists ------------------------------------------------------------------------
hid_t H5Pcreate(hid_t cls)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  Handle o; o.kind = H5I_GENPROP_LST; o.plist_class = cls;
  return new_handle(std::move(o));
}
herr_t H5Pclose(hid_t p) { return close_handle(p); }
hid_t H5Pget_driver(hid_t) { return 0; }   // never H5FD_MPIO
herr_t H5Pset_dxpl_mpio(hid_t, H5FD_mpio_xfer_t) { return 0; }

} // extern "C"
