// libevc_reader: native TFRecord / SequenceExample decoder for the frame-level YouTube-8M input path
// (include/evc_reader.h).  Plain C++17 + POSIX, no CUDA, no protobuf library: the SequenceExample wire
// format is walked by hand and every feature row is copied exactly once, from the memory-mapped shard
// into the caller's (pinned) batch buffer, by a pool of worker threads.
#include "../../include/evc_reader.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local char g_err[512] = "";
char g_err_shared[512] = "";   // last error of a worker thread, copied to the caller's thread by evc_reader_next
std::mutex g_err_mutex;

int fail(const char* fmt, const char* a = "", long long b = 0) {
  std::snprintf(g_err, sizeof(g_err), fmt, a, b);
  return -1;
}

// ---------------------------------------------------------------- CRC32C (Castagnoli), slicing-by-8
struct Crc32cTable {
  uint32_t t[8][256];
  Crc32cTable() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (0x82F63B78u & (0u - (c & 1u)));
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xFF];
  }
};
const Crc32cTable& crc_table() {
  static const Crc32cTable tab;
  return tab;
}
uint32_t crc32c(const unsigned char* p, size_t n) {
#if defined(__SSE4_2__)
  // the crc32 instruction implements exactly this polynomial
  uint64_t c64 = 0xFFFFFFFFu;
  while (n >= 8) {
    uint64_t v;
    std::memcpy(&v, p, 8);
    c64 = __builtin_ia32_crc32di(c64, v);
    p += 8;
    n -= 8;
  }
  uint32_t c32 = static_cast<uint32_t>(c64);
  while (n--) c32 = __builtin_ia32_crc32qi(c32, *p++);
  return c32 ^ 0xFFFFFFFFu;
#endif
  const Crc32cTable& T = crc_table();
  uint32_t c = 0xFFFFFFFFu;
  while (n >= 8) {
    uint32_t lo, hi;
    std::memcpy(&lo, p, 4);
    std::memcpy(&hi, p + 4, 4);
    lo ^= c;
    c = T.t[7][lo & 0xFF] ^ T.t[6][(lo >> 8) & 0xFF] ^ T.t[5][(lo >> 16) & 0xFF] ^ T.t[4][lo >> 24] ^
        T.t[3][hi & 0xFF] ^ T.t[2][(hi >> 8) & 0xFF] ^ T.t[1][(hi >> 16) & 0xFF] ^ T.t[0][hi >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) c = (c >> 8) ^ T.t[0][(c ^ *p++) & 0xFF];
  return c ^ 0xFFFFFFFFu;
}
uint32_t mask_crc(uint32_t crc) { return ((crc >> 15) | (crc << 17)) + 0xa282ead8u; }

// ---------------------------------------------------------------- protobuf wire format (just enough)
struct Span {
  const unsigned char* p = nullptr;
  size_t n = 0;
};

bool read_varint(const unsigned char*& p, const unsigned char* end, uint64_t& v) {
  v = 0;
  for (int shift = 0; shift < 64 && p < end; shift += 7) {
    const unsigned char b = *p++;
    v |= static_cast<uint64_t>(b & 0x7F) << shift;
    if (!(b & 0x80)) return true;
  }
  return false;
}

// One field of a message: number, wire type, varint value or payload span.  Returns false at the end of the
// message or on malformed input (ok == false then).
struct Field {
  uint32_t num = 0, wt = 0;
  uint64_t value = 0;
  Span data;
};
bool next_field(const unsigned char*& p, const unsigned char* end, Field& f, bool& ok) {
  if (p >= end) return false;
  uint64_t key;
  if (!read_varint(p, end, key)) { ok = false; return false; }
  f.num = static_cast<uint32_t>(key >> 3);
  f.wt = static_cast<uint32_t>(key & 7);
  switch (f.wt) {
    case 0:
      if (!read_varint(p, end, f.value)) { ok = false; return false; }
      return true;
    case 2: {
      uint64_t len;
      if (!read_varint(p, end, len) || len > static_cast<uint64_t>(end - p)) { ok = false; return false; }
      f.data = Span{p, static_cast<size_t>(len)};
      p += len;
      return true;
    }
    case 1:
      if (end - p < 8) { ok = false; return false; }
      f.data = Span{p, 8};
      p += 8;
      return true;
    case 5:
      if (end - p < 4) { ok = false; return false; }
      f.data = Span{p, 4};
      p += 4;
      return true;
    default:   // groups (3, 4) do not occur in tf.train.SequenceExample
      ok = false;
      return false;
  }
}

// map<string, X> entry: field 1 = key, field 2 = value
bool map_entry(Span entry, Span& key, Span& value) {
  const unsigned char* p = entry.p;
  const unsigned char* end = p + entry.n;
  bool ok = true;
  Field f;
  key = Span{};
  value = Span{};
  while (next_field(p, end, f, ok)) {
    if (f.wt != 2) continue;
    if (f.num == 1) key = f.data;
    else if (f.num == 2) value = f.data;
  }
  return ok;
}

bool key_is(Span key, const std::string& s) { return key.n == s.size() && std::memcmp(key.p, s.data(), key.n) == 0; }

// tensorflow.Feature { BytesList bytes_list = 1; FloatList float_list = 2; Int64List int64_list = 3; }
// first value of the bytes_list (the YT8M layout stores one bytes value per feature)
bool feature_first_bytes(Span feature, Span& out, bool& found) {
  const unsigned char* p = feature.p;
  const unsigned char* end = p + feature.n;
  bool ok = true;
  Field f;
  found = false;
  while (next_field(p, end, f, ok)) {
    if (f.num != 1 || f.wt != 2) continue;
    const unsigned char* q = f.data.p;
    const unsigned char* qe = q + f.data.n;
    Field g;
    while (next_field(q, qe, g, ok)) {
      if (g.num == 1 && g.wt == 2) {
        out = g.data;
        found = true;
        return ok;
      }
    }
    return ok;
  }
  return ok;
}

// int64_list values (packed or one varint per element) -> set labels[idx] = 1 for 0 <= idx < num_classes
bool feature_labels(Span feature, unsigned char* labels, int num_classes) {
  const unsigned char* p = feature.p;
  const unsigned char* end = p + feature.n;
  bool ok = true;
  Field f;
  auto set = [&](uint64_t raw) {
    const int64_t v = static_cast<int64_t>(raw);
    if (v >= 0 && v < num_classes) labels[v] = 1;   // sparse_to_dense(validate_indices=False): others are dropped
  };
  while (next_field(p, end, f, ok)) {
    if (f.num != 3 || f.wt != 2) continue;
    const unsigned char* q = f.data.p;
    const unsigned char* qe = q + f.data.n;
    Field g;
    while (next_field(q, qe, g, ok)) {
      if (g.num != 1) continue;
      if (g.wt == 0) {
        set(g.value);
      } else if (g.wt == 2) {
        const unsigned char* r = g.data.p;
        const unsigned char* re = r + g.data.n;
        uint64_t v;
        while (r < re) {
          if (!read_varint(r, re, v)) return false;
          set(v);
        }
      }
    }
  }
  return ok;
}

struct Mapping {
  unsigned char* base = nullptr;
  size_t size = 0;
  std::string path;
  ~Mapping() {
    if (base != nullptr && size > 0) munmap(base, size);
  }
};

struct Record {
  const unsigned char* p;
  size_t n;
  bool check_crc;   // the masked CRC32C of the payload follows it; verified by the worker that decodes it
};

struct Config {
  std::vector<std::string> names;
  std::vector<int> sizes;
  int D = 0, num_classes = 0, max_frames = 0;
};

// Decodes one SequenceExample into its slot of the batch.  Returns 0 or -1 (message in g_err).
int decode_record(const Config& cfg, Record rec, unsigned char* feat, unsigned char* labels, int* num_frames,
                  char* id, int id_stride) {
  const size_t feat_bytes = static_cast<size_t>(cfg.max_frames) * cfg.D;
  if (rec.check_crc) {
    uint32_t want;
    std::memcpy(&want, rec.p + rec.n, 4);
    if (mask_crc(crc32c(rec.p, rec.n)) != want) return fail("corrupted record data (CRC32C mismatch)%s", "");
  }
  std::memset(labels, 0, cfg.num_classes);
  if (id != nullptr && id_stride > 0) std::memset(id, 0, id_stride);
  const unsigned char* p = rec.p;
  const unsigned char* end = p + rec.n;
  bool ok = true;
  Field top;
  std::vector<int> frames(cfg.names.size(), -1);
  bool have_id = false, have_labels = false;
  while (next_field(p, end, top, ok)) {
    if (top.wt != 2) continue;
    const unsigned char* q = top.data.p;
    const unsigned char* qe = q + top.data.n;
    Field e;
    if (top.num == 1) {   // context: Features { map<string, Feature> feature = 1; }
      while (next_field(q, qe, e, ok)) {
        if (e.num != 1 || e.wt != 2) continue;
        Span key, val;
        if (!map_entry(e.data, key, val)) return fail("malformed context entry%s", "");
        if (key_is(key, "id")) {
          Span s;
          bool found;
          if (!feature_first_bytes(val, s, found)) return fail("malformed 'id' feature%s", "");
          if (found && id != nullptr && id_stride > 0)
            std::memcpy(id, s.p, s.n < static_cast<size_t>(id_stride - 1) ? s.n : static_cast<size_t>(id_stride - 1));
          have_id = found;
        } else if (key_is(key, "labels")) {
          if (!feature_labels(val, labels, cfg.num_classes)) return fail("malformed 'labels' feature%s", "");
          have_labels = true;
        }
      }
    } else if (top.num == 2) {   // feature_lists: FeatureLists { map<string, FeatureList> feature_list = 1; }
      while (next_field(q, qe, e, ok)) {
        if (e.num != 1 || e.wt != 2) continue;
        Span key, val;
        if (!map_entry(e.data, key, val)) return fail("malformed feature_lists entry%s", "");
        int which = -1, col = 0;
        for (size_t i = 0; i < cfg.names.size(); ++i) {
          if (key_is(key, cfg.names[i])) { which = static_cast<int>(i); break; }
          col += cfg.sizes[i];
        }
        if (which < 0) continue;
        const int size = cfg.sizes[which];
        // FeatureList { repeated Feature feature = 1; }: one Feature per frame
        const unsigned char* r = val.p;
        const unsigned char* re = r + val.n;
        Field fr;
        int n = 0;
        while (next_field(r, re, fr, ok)) {
          if (fr.num != 1 || fr.wt != 2) continue;
          if (n < cfg.max_frames) {
            Span row;
            bool found;
            if (!feature_first_bytes(fr.data, row, found) || !found)
              return fail("feature '%s': frame without a bytes value", cfg.names[which].c_str());
            if (row.n != static_cast<size_t>(size))
              return fail("feature '%s': a frame has %lld bytes, not the declared size", cfg.names[which].c_str(),
                          static_cast<long long>(row.n));
            std::memcpy(feat + static_cast<size_t>(n) * cfg.D + col, row.p, size);
          }
          ++n;
        }
        frames[which] = n;
      }
    }
  }
  if (!ok) return fail("malformed SequenceExample (bad varint, length or wire type)%s", "");
  if (!have_id) return fail("SequenceExample without the 'id' context feature%s", "");
  (void)have_labels;   // a VarLenFeature may be absent: no labels
  int n0 = -1;
  for (size_t i = 0; i < cfg.names.size(); ++i) {
    if (frames[i] < 0) return fail("SequenceExample without the feature list '%s'", cfg.names[i].c_str());
    if (n0 < 0) n0 = frames[i];
    else if (frames[i] != n0) return fail("feature '%s' has %lld frames, the first feature another count", cfg.names[i].c_str(), frames[i]);
  }
  const int k = n0 < cfg.max_frames ? n0 : cfg.max_frames;
  // resize_axis(feature_matrix, 0, max_frames): zero rows past the last frame
  std::memset(feat + static_cast<size_t>(k) * cfg.D, 0, feat_bytes - static_cast<size_t>(k) * cfg.D);
  *num_frames = k;
  return 0;
}

}  // namespace

struct evc_reader {
  Config cfg;
  std::vector<std::string> paths;
  int num_threads = 1;
  bool verify_crc = false;
  size_t next_path = 0;
  std::shared_ptr<Mapping> cur;
  size_t cur_off = 0;
  long long position = 0;
};

namespace {

int open_next_shard(evc_reader* r) {
  r->cur.reset();
  r->cur_off = 0;
  while (r->next_path < r->paths.size()) {
    const std::string& path = r->paths[r->next_path++];
    const int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) return fail("cannot open %s", path.c_str());
    struct stat st;
    if (fstat(fd, &st) != 0) {
      ::close(fd);
      return fail("cannot stat %s", path.c_str());
    }
    if (st.st_size == 0) {
      ::close(fd);
      continue;
    }
    void* base = mmap(nullptr, static_cast<size_t>(st.st_size), PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (base == MAP_FAILED) return fail("cannot mmap %s", path.c_str());
    madvise(base, static_cast<size_t>(st.st_size), MADV_SEQUENTIAL);
    auto m = std::make_shared<Mapping>();
    m->base = static_cast<unsigned char*>(base);
    m->size = static_cast<size_t>(st.st_size);
    m->path = path;
    r->cur = m;
    return 1;
  }
  return 0;
}

// next record of the stream: 1 = found, 0 = end of all shards, -1 = error
int next_record(evc_reader* r, Record& rec, std::vector<std::shared_ptr<Mapping>>& keep) {
  for (;;) {
    if (!r->cur) {
      const int rc = open_next_shard(r);
      if (rc <= 0) return rc;
    }
    const Mapping& m = *r->cur;
    if (r->cur_off >= m.size) {
      r->cur.reset();
      continue;
    }
    if (m.size - r->cur_off < 12) return fail("truncated record header in %s", m.path.c_str());
    const unsigned char* h = m.base + r->cur_off;
    uint64_t len;
    std::memcpy(&len, h, 8);
    if (len > m.size - r->cur_off - 12 || m.size - r->cur_off - 12 - len < 4)
      return fail("truncated record in %s", m.path.c_str());
    if (r->verify_crc) {
      uint32_t want;
      std::memcpy(&want, h + 8, 4);
      if (mask_crc(crc32c(h, 8)) != want) return fail("corrupted record length (CRC32C) in %s", m.path.c_str());
    }
    rec = Record{h + 12, static_cast<size_t>(len), r->verify_crc};
    r->cur_off += 12 + len + 4;
    if (keep.empty() || keep.back() != r->cur) keep.push_back(r->cur);
    return 1;
  }
}

}  // namespace

extern "C" int evc_reader_version(void) { return 1; }
extern "C" const char* evc_reader_last_error(void) { return g_err; }

extern "C" unsigned int evc_crc32c_masked(const unsigned char* data, long long n) {
  return mask_crc(crc32c(data, n > 0 ? static_cast<size_t>(n) : 0));
}

extern "C" evc_reader* evc_reader_open(const char* const* paths, int num_paths, const char* const* feature_names,
                                       const int* feature_sizes, int num_features, int num_classes, int max_frames,
                                       int num_threads, int verify_crc) {
  if (num_paths < 0 || (num_paths > 0 && paths == nullptr)) { fail("reader_open: bad path list%s", ""); return nullptr; }
  if (num_features <= 0 || feature_names == nullptr || feature_sizes == nullptr) {
    fail("No feature selected: feature_names is empty!%s", "");   // readers.py:141-142
    return nullptr;
  }
  if (num_classes <= 0 || max_frames <= 0) { fail("reader_open: num_classes and max_frames must be positive%s", ""); return nullptr; }
  auto* r = new evc_reader();
  for (int i = 0; i < num_features; ++i) {
    if (feature_sizes[i] <= 0 || feature_names[i] == nullptr) {
      delete r;
      fail("reader_open: bad feature description%s", "");
      return nullptr;
    }
    r->cfg.names.emplace_back(feature_names[i]);
    r->cfg.sizes.push_back(feature_sizes[i]);
    r->cfg.D += feature_sizes[i];
  }
  r->cfg.num_classes = num_classes;
  r->cfg.max_frames = max_frames;
  for (int i = 0; i < num_paths; ++i) r->paths.emplace_back(paths[i]);
  const int hw = static_cast<int>(std::thread::hardware_concurrency());
  r->num_threads = num_threads > 0 ? num_threads : (hw > 0 ? hw : 1);
  r->verify_crc = verify_crc != 0;
  return r;
}

extern "C" int evc_reader_next(evc_reader* r, int batch, unsigned char* features, unsigned char* labels,
                               int* num_frames, char* ids, int id_stride) {
  if (r == nullptr || batch < 0 || features == nullptr || labels == nullptr || num_frames == nullptr)
    return fail("reader_next: null argument%s", "");
  std::vector<Record> recs;
  std::vector<std::shared_ptr<Mapping>> keep;   // shards stay mapped until their records are decoded
  recs.reserve(batch);
  while (static_cast<int>(recs.size()) < batch) {
    Record rec;
    const int rc = next_record(r, rec, keep);
    if (rc < 0) return -1;
    if (rc == 0) break;
    recs.push_back(rec);
  }
  const int n = static_cast<int>(recs.size());
  if (n == 0) return 0;
  const size_t feat_bytes = static_cast<size_t>(r->cfg.max_frames) * r->cfg.D;
  std::atomic<int> next{0};
  std::atomic<int> failed{0};
  auto work = [&]() {
    for (;;) {
      const int i = next.fetch_add(1, std::memory_order_relaxed);
      if (i >= n || failed.load(std::memory_order_relaxed)) return;
      if (decode_record(r->cfg, recs[i], features + static_cast<size_t>(i) * feat_bytes,
                        labels + static_cast<size_t>(i) * r->cfg.num_classes, num_frames + i,
                        ids ? ids + static_cast<size_t>(i) * id_stride : nullptr, id_stride) != 0) {
        std::lock_guard<std::mutex> lock(g_err_mutex);
        if (!failed.exchange(1)) std::snprintf(g_err_shared, sizeof(g_err_shared), "record %lld: %.400s", r->position + i, g_err);
        return;
      }
    }
  };
  const int threads = r->num_threads < n ? r->num_threads : n;
  std::vector<std::thread> pool;
  pool.reserve(threads > 1 ? threads - 1 : 0);
  for (int t = 1; t < threads; ++t) pool.emplace_back(work);
  work();
  for (auto& th : pool) th.join();
  if (failed.load()) {
    std::lock_guard<std::mutex> lock(g_err_mutex);
    std::snprintf(g_err, sizeof(g_err), "%s", g_err_shared);
    return -1;
  }
  r->position += n;
  return n;
}

extern "C" long long evc_reader_position(const evc_reader* r) { return r ? r->position : -1; }

extern "C" int evc_reader_rewind(evc_reader* r) {
  if (r == nullptr) return fail("reader_rewind: null reader%s", "");
  r->cur.reset();
  r->cur_off = 0;
  r->next_path = 0;
  r->position = 0;
  return 0;
}

extern "C" void evc_reader_close(evc_reader* r) { delete r; }
