// lbm_pt_writer.cpp — the reference's on-disk snapshot format without libtorch.
//
// Every reference driver ends with `torch::save(tensor, "<name>.pt")` of a contiguous CPU fp64
// tensor ({X,Y,T}, {X,Y,9,T}, {T,2}; e.g. test/horizontal_poiseuille_test.cpp:157-160,
// test/cylinder_test.cpp:168-172, test/mrtcg_rayleigh_taylor.cpp:481-485).  torch::save(Tensor)
// writes a TorchScript archive: a ZIP of STORED entries
//     <stem>/data/0                      raw little-endian storage
//     <stem>/data.pkl                    pickle: __torch__.Module with one parameter "0" rebuilt by
//                                        torch._utils._rebuild_tensor_v2(storage, 0, sizes, strides, False, OrderedDict())
//     <stem>/code/__torch__.py           class Module(Module): __parameters__ = ["0", ]
//     <stem>/code/__torch__.py.debug_pkl, constants.pkl, version ("3"), byteorder, .data/serialization_id
// (layout read off archives written by libtorch 2.11's torch::save).
// lbm_save_pt reproduces that archive — ZIP64 records when the storage exceeds 4 GiB, storage
// aligned to 64 bytes like PyTorch's writer — so offline tooling keeps loading the files with
// torch.jit.load / torch.load (Python) or torch::load (C++).  Host-only code, no CUDA.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/lbm_b200.h"

namespace lbm
{
void set_error(const char* fmt, ...);
}

namespace
{

// ---- CRC-32 (IEEE 802.3, the ZIP polynomial), slice-by-8
struct Crc32
{
  uint32_t t[8][256];
  Crc32()
  {
    for (uint32_t i = 0; i < 256; i++)
    {
      uint32_t c = i;
      for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; i++)
      for (int s = 1; s < 8; s++) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xFF];
  }
  uint32_t run(uint32_t crc, const unsigned char* p, size_t n) const
  {
    crc = ~crc;
    while (n >= 8)
    {
      uint32_t a, b;
      std::memcpy(&a, p, 4);
      std::memcpy(&b, p + 4, 4);
      a ^= crc;
      crc = t[7][a & 0xFF] ^ t[6][(a >> 8) & 0xFF] ^ t[5][(a >> 16) & 0xFF] ^ t[4][a >> 24] ^ t[3][b & 0xFF] ^
            t[2][(b >> 8) & 0xFF] ^ t[1][(b >> 16) & 0xFF] ^ t[0][b >> 24];
      p += 8;
      n -= 8;
    }
    while (n--) crc = t[0][(crc ^ *p++) & 0xFF] ^ (crc >> 8);
    return ~crc;
  }
};

// ---- pickle protocol 2 helpers
void put_int(std::string& s, long long v)
{
  if (v >= 0 && v < 256) { s += 'K'; s += (char)v; }
  else if (v >= 0 && v < 65536) { s += 'M'; s += (char)(v & 0xFF); s += (char)(v >> 8); }
  else if (v >= INT32_MIN && v <= INT32_MAX)
  {
    s += 'J';
    const int32_t w = (int32_t)v;
    s.append(reinterpret_cast<const char*>(&w), 4);
  }
  else
  {
    s += '\x8a';  // LONG1: length byte + little-endian two's complement
    s += (char)8;
    s.append(reinterpret_cast<const char*>(&v), 8);
  }
}

void put_unicode(std::string& s, const char* text)
{
  const uint32_t n = (uint32_t)std::strlen(text);
  s += 'X';
  s.append(reinterpret_cast<const char*>(&n), 4);
  s.append(text, n);
}

std::string data_pkl(const long long* shape, int ndim, long long numel)
{
  std::string s("\x80\x02", 2);
  s += "c__torch__\nModule\nq";
  s += '\0';
  s += ")\x81}(";
  put_unicode(s, "0");
  s += "q\x01";
  s += "ctorch._utils\n_rebuild_tensor_v2\nq\x02((";
  put_unicode(s, "storage");
  s += "q\x03";
  s += "ctorch\nDoubleStorage\nq\x04h\x01";
  put_unicode(s, "cpu");
  s += "q\x05";
  put_int(s, numel);
  s += "tQq\x06K";
  s += '\0';  // storage offset 0
  s += '(';
  for (int i = 0; i < ndim; i++) put_int(s, shape[i]);
  s += "t(";
  long long stride = numel;
  for (int i = 0; i < ndim; i++)
  {
    stride = shape[i] > 0 ? stride / shape[i] : 0;
    put_int(s, stride);
  }
  s += "t\x89";
  s += "ccollections\nOrderedDict\nq\x07)RtRq\x08ubq\t.";
  return s;
}

struct Entry
{
  std::string name;
  const unsigned char* data;
  uint64_t size;
  uint32_t crc;
  uint64_t offset;  // of the local header
};

void le16(std::string& s, uint16_t v) { s.append(reinterpret_cast<const char*>(&v), 2); }
void le32(std::string& s, uint32_t v) { s.append(reinterpret_cast<const char*>(&v), 4); }
void le64(std::string& s, uint64_t v) { s.append(reinterpret_cast<const char*>(&v), 8); }

constexpr uint32_t MAX32 = 0xFFFFFFFFu;

}  // namespace

extern "C" int lbm_save_pt(const char* path, const double* data, const long long* shape, int ndim)
{
  if (!path || !shape || ndim < 0 || ndim > 16) { lbm::set_error("lbm_save_pt: bad argument"); return LBM_ERR_INVALID; }
  long long numel = 1;
  for (int i = 0; i < ndim; i++)
  {
    if (shape[i] < 0) { lbm::set_error("lbm_save_pt: negative extent"); return LBM_ERR_INVALID; }
    numel *= shape[i];
  }
  if (numel > 0 && !data) { lbm::set_error("lbm_save_pt: null data"); return LBM_ERR_INVALID; }

  // archive name = file stem, like caffe2::serialize::PyTorchStreamWriter
  std::string stem(path);
  const size_t slash = stem.find_last_of("/\\");
  if (slash != std::string::npos) stem = stem.substr(slash + 1);
  const size_t dot = stem.find_last_of('.');
  if (dot != std::string::npos && dot > 0) stem = stem.substr(0, dot);
  if (stem.empty()) stem = "archive";

  static const Crc32 crc;
  const std::string pkl = data_pkl(shape, ndim, numel);
  const std::string code = "class Module(Module):\n  __parameters__ = [\"0\", ]\n  __buffers__ = []\n  __annotations__ = []\n"
                           "  __annotations__[\"0\"] = Tensor\n";
  static const char debug_head[] = "\x80\x02X\x18\x00\x00\x00" "FORMAT_WITH_STRING_TABLEq\x00X\x00\x00\x00\x00q\x01\x85q\x02K\x00";
  std::string debug(debug_head, sizeof(debug_head) - 1);
  debug += "ctorch.jit._pickle\nbuild_intlist\nq\x03(](etRK";
  debug += '\0';
  debug += 'K';
  debug += '\0';
  debug += "\x87q\x04K";
  debug += '\0';
  debug += 'K';
  debug += '\0';
  debug += "\x87K";
  debug += '\0';
  debug += "\x87\x85q\x05\x87.";
  const std::string constants("\x80\x02).", 4);
  const std::string version = "3\n", byteorder = "little";
  char idbuf[48];
  std::snprintf(idbuf, sizeof(idbuf), "%020llu%020llu", (unsigned long long)crc.run(0, (const unsigned char*)pkl.data(), pkl.size()),
                (unsigned long long)numel);
  const std::string serial(idbuf, 40);

  auto bytes = [](const std::string& s) { return reinterpret_cast<const unsigned char*>(s.data()); };
  std::vector<Entry> entries = {
      {stem + "/data/0", reinterpret_cast<const unsigned char*>(data), (uint64_t)numel * 8, 0, 0},
      {stem + "/data.pkl", bytes(pkl), pkl.size(), 0, 0},
      {stem + "/code/__torch__.py", bytes(code), code.size(), 0, 0},
      {stem + "/code/__torch__.py.debug_pkl", bytes(debug), debug.size(), 0, 0},
      {stem + "/constants.pkl", bytes(constants), constants.size(), 0, 0},
      {stem + "/version", bytes(version), version.size(), 0, 0},
      {stem + "/byteorder", bytes(byteorder), byteorder.size(), 0, 0},
      {stem + "/.data/serialization_id", bytes(serial), serial.size(), 0, 0},
  };

  FILE* fp = std::fopen(path, "wb");
  if (!fp) { lbm::set_error("lbm_save_pt: cannot open %s for writing", path); return LBM_ERR_INVALID; }
  uint64_t pos = 0;
  bool ok = true;
  auto emit = [&](const void* p, size_t n) {
    if (n && std::fwrite(p, 1, n, fp) != n) ok = false;
    pos += n;
  };

  for (Entry& e : entries)
  {
    // CRC in chunks (the storage may be many GiB)
    uint32_t c = 0;
    for (uint64_t done = 0; done < e.size;)
    {
      const size_t n = (size_t)std::min<uint64_t>(e.size - done, 1u << 30);
      c = crc.run(c, e.data + done, n);
      done += n;
    }
    e.crc = c;
    e.offset = pos;
    const bool big = e.size >= MAX32;
    // extra field: [zip64 sizes] + "FB" padding so that the data starts on a 64-byte boundary
    std::string extra;
    if (big)
    {
      le16(extra, 0x0001);
      le16(extra, 16);
      le64(extra, e.size);
      le64(extra, e.size);
    }
    const uint64_t data_start_unpadded = pos + 30 + e.name.size() + extra.size() + 4;
    const size_t pad = (size_t)((64 - data_start_unpadded % 64) % 64);
    extra += "FB";
    le16(extra, (uint16_t)pad);
    extra.append(pad, 'Z');
    std::string h;
    le32(h, 0x04034b50);
    le16(h, big ? 45 : 20);  // version needed
    le16(h, 0x0800);         // UTF-8 names
    le16(h, 0);              // stored
    le16(h, 0);
    le16(h, 0x21);           // time / date (1980-01-01)
    le32(h, e.crc);
    le32(h, big ? MAX32 : (uint32_t)e.size);
    le32(h, big ? MAX32 : (uint32_t)e.size);
    le16(h, (uint16_t)e.name.size());
    le16(h, (uint16_t)extra.size());
    emit(h.data(), h.size());
    emit(e.name.data(), e.name.size());
    emit(extra.data(), extra.size());
    for (uint64_t done = 0; done < e.size && ok;)
    {
      const size_t n = (size_t)std::min<uint64_t>(e.size - done, 1u << 30);
      emit(e.data + done, n);
      done += n;
    }
  }

  // central directory
  const uint64_t cd_start = pos;
  bool need64 = false;
  for (const Entry& e : entries)
  {
    const bool big = e.size >= MAX32, far = e.offset >= MAX32;
    need64 = need64 || big || far;
    std::string extra;
    if (big || far)
    {
      le16(extra, 0x0001);
      le16(extra, (uint16_t)((big ? 16 : 0) + (far ? 8 : 0)));
      if (big) { le64(extra, e.size); le64(extra, e.size); }
      if (far) le64(extra, e.offset);
    }
    std::string h;
    le32(h, 0x02014b50);
    le16(h, 45);             // version made by
    le16(h, (big || far) ? 45 : 20);
    le16(h, 0x0800);
    le16(h, 0);
    le16(h, 0);
    le16(h, 0x21);
    le32(h, e.crc);
    le32(h, big ? MAX32 : (uint32_t)e.size);
    le32(h, big ? MAX32 : (uint32_t)e.size);
    le16(h, (uint16_t)e.name.size());
    le16(h, (uint16_t)extra.size());
    le16(h, 0);              // comment
    le16(h, 0);              // disk
    le16(h, 0);              // internal attributes
    le32(h, 0);              // external attributes
    le32(h, far ? MAX32 : (uint32_t)e.offset);
    emit(h.data(), h.size());
    emit(e.name.data(), e.name.size());
    emit(extra.data(), extra.size());
  }
  const uint64_t cd_size = pos - cd_start;
  need64 = need64 || cd_start >= MAX32;
  if (need64)
  {
    const uint64_t z64 = pos;
    std::string r;
    le32(r, 0x06064b50);
    le64(r, 44);
    le16(r, 45);
    le16(r, 45);
    le32(r, 0);
    le32(r, 0);
    le64(r, entries.size());
    le64(r, entries.size());
    le64(r, cd_size);
    le64(r, cd_start);
    le32(r, 0x07064b50);  // locator
    le32(r, 0);
    le64(r, z64);
    le32(r, 1);
    emit(r.data(), r.size());
  }
  std::string eocd;
  le32(eocd, 0x06054b50);
  le16(eocd, 0);
  le16(eocd, 0);
  le16(eocd, (uint16_t)entries.size());
  le16(eocd, (uint16_t)entries.size());
  le32(eocd, cd_size >= MAX32 ? MAX32 : (uint32_t)cd_size);
  le32(eocd, cd_start >= MAX32 ? MAX32 : (uint32_t)cd_start);
  le16(eocd, 0);
  emit(eocd.data(), eocd.size());
  if (std::fclose(fp) != 0) ok = false;
  if (!ok) { lbm::set_error("lbm_save_pt: write to %s failed", path); return LBM_ERR_INVALID; }
  return LBM_OK;
}
