// LASzip codec (host code; file decoding / encoding sits next to the hot path).  The reference
// gets this from laspy[lazrs] (src/vapc/datahandler.py:72-73, :99-165), which is not installable offline.  Written from the
// published LASzip format description (Isenburg, "LASzip: lossless compression of LiDAR data", PE&RS 2013): an adaptive
// arithmetic coder (Said's FastAC scheme), integer "corrector" coders with k-bit contexts and the per-item predictors.
//   compressor 2, "pointwise chunked":  POINT10 v2, GPSTIME11 v2, RGB12 v2, BYTE v2          LAS point formats 0 - 3
//   compressor 3, "layered chunked":    POINT14 v3, RGB14 v3, BYTE14 v3                       LAS 1.4 point formats 6 / 7
// Decoder pins: the reference's four fixtures decode to their headers' bounding boxes and by-return counts, and every chunk /
// layer decoder ends on the last byte of its stream (enforced for every file).  Encoder pin: re-encoding the decoded records
// reproduces the point data of all four files byte for byte (tests/test_laz.py).  GPSTIME11 v2 and RGB14 v3 occur in none of
// the fixtures: they reuse pinned coders (the GPS coder of POINT14 v3, the colour coder of RGB12 v2) and are round-trip tested.
// Chunks are independent: they are decoded / encoded on all host cores.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {
// Chunks are independent streams: decode / encode them on the host's cores.  make() builds one set of coder objects per thread,
// body(chunk, state) returns a status whose message lives in the worker's thread-local slot; the first failure wins and is
// re-raised in the caller's thread.  VAPC_LAZ_THREADS overrides the thread count.
template <typename Make, typename Body>
int parallel_chunks(int64_t nchunks, Make make, Body body) {
  unsigned hw = std::thread::hardware_concurrency();
  if (const char* env = getenv("VAPC_LAZ_THREADS")) hw = (unsigned)atoi(env);
  if (hw < 1) hw = 1;
  const int64_t nthreads = std::min<int64_t>(std::min<int64_t>(hw, 32), nchunks);
  std::atomic<int64_t> next{0};
  std::atomic<int> failed{0};
  std::mutex mu;
  std::string msg;
  int code = VAPC_OK;
  auto worker = [&]() {
    auto st = make();
    for (;;) {
      const int64_t c = next.fetch_add(1);
      if (c >= nchunks || failed.load()) break;
      const int rc = body(c, *st);
      if (rc != VAPC_OK) {
        std::lock_guard<std::mutex> lk(mu);
        if (!failed.load()) {
          code = rc;
          msg = vapc_last_error();
          failed.store(1);
        }
        break;
      }
    }
  };
  if (nthreads <= 1) {
    worker();
  } else {
    std::vector<std::thread> pool;
    for (int64_t t = 0; t < nthreads; ++t) pool.emplace_back(worker);
    for (auto& t : pool) t.join();
  }
  if (failed.load()) return vapc_set_error(code, "%s", msg.c_str());
  return VAPC_OK;
}

// ---- byte source --------------------------------------------------------------------------------------------
struct ByteReader {
  const uint8_t* p;
  const uint8_t* end;
  bool overrun = false;
  uint8_t byte() {
    if (p < end) return *p++;
    overrun = true;
    return 0;
  }
};

// ---- adaptive arithmetic decoder ------------------------------------------------------------------------------
constexpr uint32_t kAcMinLength = 0x01000000u, kAcMaxLength = 0xFFFFFFFFu;
constexpr int kBmLengthShift = 13, kDmLengthShift = 15;
constexpr uint32_t kBmMaxCount = 1u << kBmLengthShift, kDmMaxCount = 1u << kDmLengthShift;

struct BitModel {
  uint32_t update_cycle, bits_until_update, bit_0_prob, bit_0_count, bit_count;
  void init() {
    bit_0_count = 1;
    bit_count = 2;
    bit_0_prob = 1u << (kBmLengthShift - 1);
    update_cycle = bits_until_update = 4;
  }
  void update() {
    if ((bit_count += update_cycle) > kBmMaxCount) {
      bit_count = (bit_count + 1) >> 1;
      bit_0_count = (bit_0_count + 1) >> 1;
      if (bit_0_count == bit_count) ++bit_count;
    }
    const uint32_t scale = 0x80000000u / bit_count;
    bit_0_prob = (bit_0_count * scale) >> (31 - kBmLengthShift);
    update_cycle = (5 * update_cycle) >> 2;
    if (update_cycle > 64) update_cycle = 64;
    bits_until_update = update_cycle;
  }
};

struct SymbolModel {
  uint32_t symbols = 0, last_symbol = 0, table_size = 0, table_shift = 0;
  uint32_t total_count = 0, update_cycle = 0, symbols_until_update = 0;
  std::vector<uint32_t> distribution, symbol_count, decoder_table;
  explicit SymbolModel(uint32_t n) : symbols(n) {
    last_symbol = n - 1;
    if (n > 16) {
      uint32_t table_bits = 3;
      while (n > (1u << (table_bits + 2))) ++table_bits;
      table_size = 1u << table_bits;
      table_shift = kDmLengthShift - table_bits;
      decoder_table.assign(table_size + 2, 0);
    }
    distribution.assign(n, 0);
    symbol_count.assign(n, 0);
    init();
  }
  void init() {
    total_count = 0;
    update_cycle = symbols;
    for (uint32_t k = 0; k < symbols; ++k) symbol_count[k] = 1;
    update();
    symbols_until_update = update_cycle = (symbols + 6) >> 1;
  }
  void update() {
    if ((total_count += update_cycle) > kDmMaxCount) {
      total_count = 0;
      for (uint32_t n = 0; n < symbols; ++n) total_count += (symbol_count[n] = (symbol_count[n] + 1) >> 1);
    }
    uint32_t sum = 0, s = 0;
    const uint32_t scale = 0x80000000u / total_count;
    if (table_size == 0) {
      for (uint32_t k = 0; k < symbols; ++k) {
        distribution[k] = (scale * sum) >> (31 - kDmLengthShift);
        sum += symbol_count[k];
      }
    } else {
      for (uint32_t k = 0; k < symbols; ++k) {
        distribution[k] = (scale * sum) >> (31 - kDmLengthShift);
        sum += symbol_count[k];
        const uint32_t w = distribution[k] >> table_shift;
        while (s < w) decoder_table[++s] = k - 1;
      }
      decoder_table[0] = 0;
      while (s <= table_size) decoder_table[++s] = symbols - 1;
    }
    update_cycle = (5 * update_cycle) >> 2;
    const uint32_t max_cycle = (symbols + 6) << 3;
    if (update_cycle > max_cycle) update_cycle = max_cycle;
    symbols_until_update = update_cycle;
  }
};

struct ArithmeticDecoder {
  ByteReader* in = nullptr;
  uint32_t value = 0, length = 0;
  void init(ByteReader* src) {
    in = src;
    length = kAcMaxLength;
    value = (uint32_t)in->byte() << 24;
    value |= (uint32_t)in->byte() << 16;
    value |= (uint32_t)in->byte() << 8;
    value |= (uint32_t)in->byte();
  }
  void renorm() {
    do {
      value = (value << 8) | in->byte();
    } while ((length <<= 8) < kAcMinLength);
  }
  uint32_t bit(BitModel& m) {
    const uint32_t x = m.bit_0_prob * (length >> kBmLengthShift);
    const uint32_t sym = value >= x;
    if (sym == 0) {
      length = x;
      ++m.bit_0_count;
    } else {
      value -= x;
      length -= x;
    }
    if (length < kAcMinLength) renorm();
    if (--m.bits_until_update == 0) m.update();
    return sym;
  }
  uint32_t symbol(SymbolModel& m) {
    uint32_t n, sym, x, y = length;
    if (m.table_size) {
      const uint32_t dv = value / (length >>= kDmLengthShift);
      const uint32_t t = dv >> m.table_shift;
      sym = m.decoder_table[t];
      n = m.decoder_table[t + 1] + 1;
      while (n > sym + 1) {
        const uint32_t k = (sym + n) >> 1;
        if (m.distribution[k] > dv) n = k; else sym = k;
      }
      x = m.distribution[sym] * length;
      if (sym != m.last_symbol) y = m.distribution[sym + 1] * length;
    } else {
      x = sym = 0;
      length >>= kDmLengthShift;
      uint32_t k = (n = m.symbols) >> 1;
      do {
        const uint32_t z = length * m.distribution[k];
        if (z > value) {
          n = k;
          y = z;
        } else {
          sym = k;
          x = z;
        }
      } while ((k = (sym + n) >> 1) != sym);
    }
    value -= x;
    length = y - x;
    if (length < kAcMinLength) renorm();
    ++m.symbol_count[sym];
    if (--m.symbols_until_update == 0) m.update();
    return sym;
  }
  uint32_t read_short() {
    const uint32_t sym = value / (length >>= 16);
    value -= length * sym;
    if (length < kAcMinLength) renorm();
    return sym;
  }
  uint32_t read_bits(uint32_t bits) {
    if (bits > 19) {
      const uint32_t lower = read_short();
      return (read_bits(bits - 16) << 16) | lower;
    }
    const uint32_t sym = value / (length >>= bits);
    value -= length * sym;
    if (length < kAcMinLength) renorm();
    return sym;
  }
};

// ---- integer corrector coder -------------------------------------------------------------------------------------
struct IntegerDecompressor {
  ArithmeticDecoder* dec;
  uint32_t k = 0, contexts, bits_high = 8, corr_bits, corr_range;
  int32_t corr_min;
  std::vector<std::unique_ptr<SymbolModel>> m_bits, m_corrector;  // m_corrector[0] unused: the 1-bit case is a BitModel
  BitModel corrector0;
  IntegerDecompressor(ArithmeticDecoder* d, uint32_t bits, uint32_t ctx) : dec(d), contexts(ctx) {
    if (bits && bits < 32) {
      corr_bits = bits;
      corr_range = 1u << bits;
      corr_min = -(int32_t)(corr_range / 2);
    } else {
      corr_bits = 32;
      corr_range = 0;
      corr_min = INT32_MIN;
    }
    for (uint32_t i = 0; i < contexts; ++i) m_bits.emplace_back(new SymbolModel(corr_bits + 1));
    m_corrector.emplace_back(nullptr);
    for (uint32_t i = 1; i <= corr_bits; ++i) m_corrector.emplace_back(new SymbolModel(i <= bits_high ? (1u << i) : (1u << bits_high)));
    init();
  }
  void init() {
    for (auto& m : m_bits) m->init();
    corrector0.init();
    for (uint32_t i = 1; i <= corr_bits; ++i) m_corrector[i]->init();
  }
  int32_t corrector(SymbolModel& bits_model) {
    int32_t c;
    k = dec->symbol(bits_model);
    if (k) {
      if (k < 32) {
        if (k <= bits_high) {
          c = (int32_t)dec->symbol(*m_corrector[k]);
        } else {
          const uint32_t k1 = k - bits_high;
          c = (int32_t)dec->symbol(*m_corrector[k]);
          const int32_t c1 = (int32_t)dec->read_bits(k1);
          c = (int32_t)(((uint32_t)c << k1) | (uint32_t)c1);
        }
        if (c >= (1 << (k - 1))) c += 1; else c -= ((1 << k) - 1);
      } else {
        c = corr_min;
      }
    } else {
      c = (int32_t)dec->bit(corrector0);
    }
    return c;
  }
  int32_t decompress(int32_t pred, uint32_t context) {
    int32_t real = (int32_t)((uint32_t)pred + (uint32_t)corrector(*m_bits[context]));
    if (real < 0) real = (int32_t)((uint32_t)real + corr_range);
    else if ((uint32_t)real >= corr_range) real = (int32_t)((uint32_t)real - corr_range);
    return real;
  }
};

// ---- streaming median of the last five values ----------------------------------------------------------------------
struct Median5 {
  int32_t v[5];
  bool high;
  void init() {
    v[0] = v[1] = v[2] = v[3] = v[4] = 0;
    high = true;
  }
  int32_t get() const { return v[2]; }
  void add(int32_t x) {
    if (high) {
      if (x < v[2]) {
        v[4] = v[3];
        v[3] = v[2];
        if (x < v[0]) { v[2] = v[1]; v[1] = v[0]; v[0] = x; }
        else if (x < v[1]) { v[2] = v[1]; v[1] = x; }
        else v[2] = x;
      } else {
        if (x < v[3]) { v[4] = v[3]; v[3] = x; } else v[4] = x;
        high = false;
      }
    } else {
      if (v[2] < x) {
        v[0] = v[1];
        v[1] = v[2];
        if (v[4] < x) { v[2] = v[3]; v[3] = v[4]; v[4] = x; }
        else if (v[3] < x) { v[2] = v[3]; v[3] = x; }
        else v[2] = x;
      } else {
        if (v[1] < x) { v[0] = v[1]; v[1] = x; } else v[0] = x;
        high = true;
      }
    }
  }
};

// return-number context tables of LASzip version 2 (indexed [number_of_returns][return_number])
const uint8_t kReturnMap[8][8] = {{15, 14, 13, 12, 11, 10, 9, 8}, {14, 0, 1, 3, 6, 10, 10, 9},  {13, 1, 2, 4, 7, 11, 11, 10}, {12, 3, 4, 5, 8, 12, 12, 11},
                                  {11, 6, 7, 8, 9, 13, 13, 12},   {10, 10, 11, 12, 13, 14, 14, 13}, {9, 10, 11, 12, 13, 14, 15, 14}, {8, 9, 10, 11, 12, 13, 14, 15}};
const uint8_t kReturnLevel[8][8] = {{0, 1, 2, 3, 4, 5, 6, 7}, {1, 0, 1, 2, 3, 4, 5, 6}, {2, 1, 0, 1, 2, 3, 4, 5}, {3, 2, 1, 0, 1, 2, 3, 4},
                                    {4, 3, 2, 1, 0, 1, 2, 3}, {5, 4, 3, 2, 1, 0, 1, 2}, {6, 5, 4, 3, 2, 1, 0, 1}, {7, 6, 5, 4, 3, 2, 1, 0}};

inline int32_t rd_i32(const uint8_t* p) { int32_t v; memcpy(&v, p, 4); return v; }
inline uint16_t rd_u16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }
inline void wr_i32(uint8_t* p, int32_t v) { memcpy(p, &v, 4); }
inline void wr_u16(uint8_t* p, uint16_t v) { memcpy(p, &v, 2); }
inline uint8_t u8_fold(int v) { return (uint8_t)(v & 0xFF); }
inline int u8_clamp(int v) { return v <= 0 ? 0 : (v >= 255 ? 255 : v); }

struct ItemReader {
  virtual ~ItemReader() {}
  virtual void init(const uint8_t* first) = 0;
  virtual void read(uint8_t* out) = 0;
};

// POINT10 version 2: X, Y, Z, intensity, return byte, classification, scan angle rank, user data, point source id
struct Point10V2 : ItemReader {
  ArithmeticDecoder* dec;
  uint8_t last[20];
  uint16_t last_intensity[16];
  Median5 mx[16], my[16];
  int32_t last_height[8];
  SymbolModel changed{64};
  SymbolModel scan_angle[2] = {SymbolModel(256), SymbolModel(256)};
  std::unique_ptr<SymbolModel> bit_byte[256], classification[256], user_data[256];
  IntegerDecompressor ic_intensity, ic_source, ic_dx, ic_dy, ic_z;
  explicit Point10V2(ArithmeticDecoder* d) : dec(d), ic_intensity(d, 16, 4), ic_source(d, 16, 1), ic_dx(d, 32, 2), ic_dy(d, 32, 22), ic_z(d, 32, 20) {}
  void init(const uint8_t* first) override {
    for (int i = 0; i < 16; ++i) {
      mx[i].init();
      my[i].init();
      last_intensity[i] = 0;
      last_height[i / 2] = 0;
    }
    changed.init();
    ic_intensity.init();
    scan_angle[0].init();
    scan_angle[1].init();
    ic_source.init();
    for (int i = 0; i < 256; ++i) {
      if (bit_byte[i]) bit_byte[i]->init();
      if (classification[i]) classification[i]->init();
      if (user_data[i]) user_data[i]->init();
    }
    ic_dx.init();
    ic_dy.init();
    ic_z.init();
    memcpy(last, first, 20);
    last[12] = last[13] = 0;  // the intensity context starts from zero
  }
  static SymbolModel& lazy(std::unique_ptr<SymbolModel>& m) {
    if (!m) m.reset(new SymbolModel(256));
    return *m;
  }
  void read(uint8_t* out) override {
    uint32_t r, n, m, l;
    const uint32_t ch = dec->symbol(changed);
    if (ch) {
      if (ch & 32) last[14] = (uint8_t)dec->symbol(lazy(bit_byte[last[14]]));
      r = last[14] & 7;
      n = (last[14] >> 3) & 7;
      m = kReturnMap[n][r];
      l = kReturnLevel[n][r];
      if (ch & 16) {
        const uint16_t v = (uint16_t)ic_intensity.decompress(last_intensity[m], m < 3 ? m : 3);
        wr_u16(last + 12, v);
        last_intensity[m] = v;
      } else {
        wr_u16(last + 12, last_intensity[m]);
      }
      if (ch & 8) last[15] = (uint8_t)dec->symbol(lazy(classification[last[15]]));
      if (ch & 4) {
        const int val = (int)dec->symbol(scan_angle[(last[14] >> 6) & 1]);
        last[16] = u8_fold(val + last[16]);
      }
      if (ch & 2) last[17] = (uint8_t)dec->symbol(lazy(user_data[last[17]]));
      if (ch & 1) wr_u16(last + 18, (uint16_t)ic_source.decompress(rd_u16(last + 18), 0));
    } else {
      r = last[14] & 7;
      n = (last[14] >> 3) & 7;
      m = kReturnMap[n][r];
      l = kReturnLevel[n][r];
    }
    // x, y from the median of the last five differences of the same return type; z from the last height of its level
    int32_t diff = ic_dx.decompress(mx[m].get(), n == 1);
    wr_i32(last, (int32_t)((uint32_t)rd_i32(last) + (uint32_t)diff));
    mx[m].add(diff);
    uint32_t kb = ic_dx.k;
    diff = ic_dy.decompress(my[m].get(), (n == 1) + (kb < 20 ? (kb & ~1u) : 20));
    wr_i32(last + 4, (int32_t)((uint32_t)rd_i32(last + 4) + (uint32_t)diff));
    my[m].add(diff);
    kb = (ic_dx.k + ic_dy.k) / 2;
    const int32_t zz = ic_z.decompress(last_height[l], (n == 1) + (kb < 18 ? (kb & ~1u) : 18));
    wr_i32(last + 8, zz);
    last_height[l] = zz;
    memcpy(out, last, 20);
  }
};

// RGB12 version 2: three 16-bit colours, coded byte-wise as differences; green and blue predicted from the red change
struct Rgb12V2 : ItemReader {
  ArithmeticDecoder* dec;
  uint16_t last[3];
  SymbolModel used{128};
  SymbolModel d0{256}, d1{256}, d2{256}, d3{256}, d4{256}, d5{256};
  explicit Rgb12V2(ArithmeticDecoder* d) : dec(d) {}
  void init(const uint8_t* first) override {
    used.init();
    d0.init(); d1.init(); d2.init(); d3.init(); d4.init(); d5.init();
    memcpy(last, first, 6);
  }
  void read(uint8_t* out) override {
    uint16_t c[3];
    int corr, diff;
    const uint32_t sym = dec->symbol(used);
    if (sym & 1) {
      corr = (int)dec->symbol(d0);
      c[0] = u8_fold(corr + (last[0] & 255));
    } else {
      c[0] = last[0] & 0xFF;
    }
    if (sym & 2) {
      corr = (int)dec->symbol(d1);
      c[0] |= (uint16_t)u8_fold(corr + (last[0] >> 8)) << 8;
    } else {
      c[0] |= last[0] & 0xFF00;
    }
    if (sym & 64) {
      diff = (c[0] & 0xFF) - (last[0] & 0xFF);
      if (sym & 4) {
        corr = (int)dec->symbol(d2);
        c[1] = u8_fold(corr + u8_clamp(diff + (last[1] & 255)));
      } else {
        c[1] = last[1] & 0xFF;
      }
      if (sym & 16) {
        corr = (int)dec->symbol(d4);
        diff = (diff + ((c[1] & 0xFF) - (last[1] & 0xFF))) / 2;
        c[2] = u8_fold(corr + u8_clamp(diff + (last[2] & 255)));
      } else {
        c[2] = last[2] & 0xFF;
      }
      diff = (c[0] >> 8) - (last[0] >> 8);
      if (sym & 8) {
        corr = (int)dec->symbol(d3);
        c[1] |= (uint16_t)u8_fold(corr + u8_clamp(diff + (last[1] >> 8))) << 8;
      } else {
        c[1] |= last[1] & 0xFF00;
      }
      if (sym & 32) {
        corr = (int)dec->symbol(d5);
        diff = (diff + ((c[1] >> 8) - (last[1] >> 8))) / 2;
        c[2] |= (uint16_t)u8_fold(corr + u8_clamp(diff + (last[2] >> 8))) << 8;
      } else {
        c[2] |= last[2] & 0xFF00;
      }
    } else {
      c[1] = c[0];
      c[2] = c[0];
    }
    memcpy(last, c, 6);
    memcpy(out, c, 6);
  }
};

// BYTE version 2 (extra bytes): every byte coded as the difference to the same byte of the previous point
struct ByteV2 : ItemReader {
  ArithmeticDecoder* dec;
  int number;
  std::vector<uint8_t> last;
  std::vector<std::unique_ptr<SymbolModel>> m_byte;
  ByteV2(ArithmeticDecoder* d, int n) : dec(d), number(n), last((size_t)n) {
    for (int i = 0; i < n; ++i) m_byte.emplace_back(new SymbolModel(256));
  }
  void init(const uint8_t* first) override {
    for (auto& m : m_byte) m->init();
    memcpy(last.data(), first, (size_t)number);
  }
  void read(uint8_t* out) override {
    for (int i = 0; i < number; ++i) last[(size_t)i] = u8_fold((int)last[(size_t)i] + (int)dec->symbol(*m_byte[(size_t)i]));
    memcpy(out, last.data(), (size_t)number);
  }
};

// ================================================================================================================
// Encoder: the exact mirror of the decoder above (same models, same contexts, same update schedule), so that the
// streams it writes are the ones LASzip writes: tests/test_laz.py re-encodes the decoded records of the reference's
// vapc_in.laz and compares the result with the file byte for byte, chunk table included.
// ================================================================================================================
struct ArithmeticEncoder {
  std::vector<uint8_t>* out = nullptr;
  size_t begin = 0;  // first byte of this coder's stream inside *out (carries never travel in front of it)
  uint32_t base = 0, length = kAcMaxLength;
  void init(std::vector<uint8_t>* dst) {
    out = dst;
    begin = dst->size();
    base = 0;
    length = kAcMaxLength;
  }
  void propagate_carry() {
    size_t i = out->size();
    while (i > begin && (*out)[i - 1] == 0xFFu) (*out)[--i] = 0;
    if (i > begin) ++(*out)[i - 1];
  }
  void renorm() {
    do {
      out->push_back((uint8_t)(base >> 24));
      base <<= 8;
    } while ((length <<= 8) < kAcMinLength);
  }
  void bit(BitModel& m, uint32_t sym) {
    const uint32_t x = m.bit_0_prob * (length >> kBmLengthShift);
    if (sym == 0) {
      length = x;
      ++m.bit_0_count;
    } else {
      const uint32_t init_base = base;
      base += x;
      length -= x;
      if (init_base > base) propagate_carry();
    }
    if (length < kAcMinLength) renorm();
    if (--m.bits_until_update == 0) m.update();
  }
  void symbol(SymbolModel& m, uint32_t sym) {
    uint32_t x;
    const uint32_t init_base = base;
    if (sym == m.last_symbol) {
      x = m.distribution[sym] * (length >> kDmLengthShift);
      base += x;
      length -= x;
    } else {
      x = m.distribution[sym] * (length >>= kDmLengthShift);
      base += x;
      length = m.distribution[sym + 1] * length - x;
    }
    if (init_base > base) propagate_carry();
    if (length < kAcMinLength) renorm();
    ++m.symbol_count[sym];
    if (--m.symbols_until_update == 0) m.update();
  }
  void write_short(uint32_t sym) {
    const uint32_t init_base = base;
    base += sym * (length >>= 16);
    if (init_base > base) propagate_carry();
    if (length < kAcMinLength) renorm();
  }
  void write_bits(uint32_t bits, uint32_t sym) {
    if (bits > 19) {
      write_short(sym & 0xFFFFu);
      sym >>= 16;
      bits -= 16;
    }
    const uint32_t init_base = base;
    base += sym * (length >>= bits);
    if (init_base > base) propagate_carry();
    if (length < kAcMinLength) renorm();
  }
  void done() {
    const uint32_t init_base = base;
    bool another_byte = true;
    if (length > 2 * kAcMinLength) {
      base += kAcMinLength;
      length = kAcMinLength >> 1;
    } else {
      base += kAcMinLength >> 1;
      length = kAcMinLength >> 9;
      another_byte = false;
    }
    if (init_base > base) propagate_carry();
    renorm();
    out->push_back(0);  // the decoder reads ahead: keep its byte reads inside the stream
    out->push_back(0);
    if (another_byte) out->push_back(0);
  }
};

struct IntegerCompressor {
  ArithmeticEncoder* enc;
  uint32_t k = 0, contexts, bits_high = 8, corr_bits, corr_range;
  int32_t corr_min, corr_max;
  std::vector<std::unique_ptr<SymbolModel>> m_bits, m_corrector;
  BitModel corrector0;
  IntegerCompressor(ArithmeticEncoder* e, uint32_t bits, uint32_t ctx) : enc(e), contexts(ctx) {
    if (bits && bits < 32) {
      corr_bits = bits;
      corr_range = 1u << bits;
      corr_min = -(int32_t)(corr_range / 2);
      corr_max = corr_min + (int32_t)corr_range - 1;
    } else {
      corr_bits = 32;
      corr_range = 0;
      corr_min = INT32_MIN;
      corr_max = INT32_MAX;
    }
    for (uint32_t i = 0; i < contexts; ++i) m_bits.emplace_back(new SymbolModel(corr_bits + 1));
    m_corrector.emplace_back(nullptr);
    for (uint32_t i = 1; i <= corr_bits; ++i) m_corrector.emplace_back(new SymbolModel(i <= bits_high ? (1u << i) : (1u << bits_high)));
    init();
  }
  void init() {
    for (auto& m : m_bits) m->init();
    corrector0.init();
    for (uint32_t i = 1; i <= corr_bits; ++i) m_corrector[i]->init();
  }
  void corrector(int32_t c, SymbolModel& bits_model) {
    uint32_t c1 = c <= 0 ? (uint32_t)0 - (uint32_t)c : (uint32_t)c - 1u;  // the tightest [-(2^k - 1), 2^k] holding c
    for (k = 0; c1; ++k) c1 >>= 1;
    enc->symbol(bits_model, k);
    if (k) {
      if (k < 32) {
        if (c < 0) c += (int32_t)((1u << k) - 1u); else c -= 1;  // -> [0, 2^k - 1]
        if (k <= bits_high) {
          enc->symbol(*m_corrector[k], (uint32_t)c);
        } else {
          const uint32_t k1 = k - bits_high;
          const uint32_t low = (uint32_t)c & ((1u << k1) - 1u);
          enc->symbol(*m_corrector[k], (uint32_t)c >> k1);
          enc->write_bits(k1, low);
        }
      }
    } else {
      enc->bit(corrector0, (uint32_t)c);
    }
  }
  void compress(int32_t pred, int32_t real, uint32_t context) {
    int32_t corr = (int32_t)((uint32_t)real - (uint32_t)pred);
    if (corr < corr_min) corr = (int32_t)((uint32_t)corr + corr_range);
    else if (corr > corr_max) corr = (int32_t)((uint32_t)corr - corr_range);
    corrector(corr, *m_bits[context]);
  }
};

struct ItemWriter {
  virtual ~ItemWriter() {}
  virtual void init(const uint8_t* first) = 0;
  virtual void write(const uint8_t* item) = 0;
};

struct Point10V2Writer : ItemWriter {
  ArithmeticEncoder* enc;
  uint8_t last[20];
  uint16_t last_intensity[16];
  Median5 mx[16], my[16];
  int32_t last_height[8];
  SymbolModel changed{64};
  SymbolModel scan_angle[2] = {SymbolModel(256), SymbolModel(256)};
  std::unique_ptr<SymbolModel> bit_byte[256], classification[256], user_data[256];
  IntegerCompressor ic_intensity, ic_source, ic_dx, ic_dy, ic_z;
  explicit Point10V2Writer(ArithmeticEncoder* e) : enc(e), ic_intensity(e, 16, 4), ic_source(e, 16, 1), ic_dx(e, 32, 2), ic_dy(e, 32, 22), ic_z(e, 32, 20) {}
  void init(const uint8_t* first) override {
    for (int i = 0; i < 16; ++i) {
      mx[i].init();
      my[i].init();
      last_intensity[i] = 0;
      last_height[i / 2] = 0;
    }
    changed.init();
    ic_intensity.init();
    scan_angle[0].init();
    scan_angle[1].init();
    ic_source.init();
    for (int i = 0; i < 256; ++i) {
      if (bit_byte[i]) bit_byte[i]->init();
      if (classification[i]) classification[i]->init();
      if (user_data[i]) user_data[i]->init();
    }
    ic_dx.init();
    ic_dy.init();
    ic_z.init();
    memcpy(last, first, 20);
    last[12] = last[13] = 0;
  }
  void write(const uint8_t* item) override {
    const uint32_t r = item[14] & 7, n = (item[14] >> 3) & 7;
    const uint32_t m = kReturnMap[n][r], l = kReturnLevel[n][r];
    const uint16_t intensity = rd_u16(item + 12);
    const uint32_t ch = ((uint32_t)(last[14] != item[14]) << 5) | ((uint32_t)(last_intensity[m] != intensity) << 4) | ((uint32_t)(last[15] != item[15]) << 3) |
                        ((uint32_t)(last[16] != item[16]) << 2) | ((uint32_t)(last[17] != item[17]) << 1) | (uint32_t)(rd_u16(last + 18) != rd_u16(item + 18));
    enc->symbol(changed, ch);
    if (ch & 32) enc->symbol(Point10V2::lazy(bit_byte[last[14]]), item[14]);
    if (ch & 16) {
      ic_intensity.compress(last_intensity[m], intensity, m < 3 ? m : 3);
      last_intensity[m] = intensity;
    }
    if (ch & 8) enc->symbol(Point10V2::lazy(classification[last[15]]), item[15]);
    if (ch & 4) enc->symbol(scan_angle[(item[14] >> 6) & 1], u8_fold((int)item[16] - (int)last[16]));
    if (ch & 2) enc->symbol(Point10V2::lazy(user_data[last[17]]), item[17]);
    if (ch & 1) ic_source.compress(rd_u16(last + 18), rd_u16(item + 18), 0);
    int32_t diff = (int32_t)((uint32_t)rd_i32(item) - (uint32_t)rd_i32(last));
    ic_dx.compress(mx[m].get(), diff, n == 1);
    mx[m].add(diff);
    uint32_t kb = ic_dx.k;
    diff = (int32_t)((uint32_t)rd_i32(item + 4) - (uint32_t)rd_i32(last + 4));
    ic_dy.compress(my[m].get(), diff, (n == 1) + (kb < 20 ? (kb & ~1u) : 20));
    my[m].add(diff);
    kb = (ic_dx.k + ic_dy.k) / 2;
    ic_z.compress(last_height[l], rd_i32(item + 8), (n == 1) + (kb < 18 ? (kb & ~1u) : 18));
    last_height[l] = rd_i32(item + 8);
    memcpy(last, item, 20);
  }
};

struct Rgb12V2Writer : ItemWriter {
  ArithmeticEncoder* enc;
  uint16_t last[3];
  SymbolModel used{128};
  SymbolModel d0{256}, d1{256}, d2{256}, d3{256}, d4{256}, d5{256};
  explicit Rgb12V2Writer(ArithmeticEncoder* e) : enc(e) {}
  void init(const uint8_t* first) override {
    used.init();
    d0.init(); d1.init(); d2.init(); d3.init(); d4.init(); d5.init();
    memcpy(last, first, 6);
  }
  void write(const uint8_t* item) override {
    uint16_t c[3];
    memcpy(c, item, 6);
    int diff_l = 0, diff_h = 0, corr;
    const uint32_t sym = ((uint32_t)((last[0] & 0x00FF) != (c[0] & 0x00FF)) << 0) | ((uint32_t)((last[0] & 0xFF00) != (c[0] & 0xFF00)) << 1) |
                         ((uint32_t)((last[1] & 0x00FF) != (c[1] & 0x00FF)) << 2) | ((uint32_t)((last[1] & 0xFF00) != (c[1] & 0xFF00)) << 3) |
                         ((uint32_t)((last[2] & 0x00FF) != (c[2] & 0x00FF)) << 4) | ((uint32_t)((last[2] & 0xFF00) != (c[2] & 0xFF00)) << 5) |
                         ((uint32_t)(((c[0] & 0x00FF) != (c[1] & 0x00FF)) || ((c[0] & 0x00FF) != (c[2] & 0x00FF)) || ((c[0] & 0xFF00) != (c[1] & 0xFF00)) ||
                                     ((c[0] & 0xFF00) != (c[2] & 0xFF00))) << 6);
    enc->symbol(used, sym);
    if (sym & 1) {
      diff_l = (int)(c[0] & 255) - (int)(last[0] & 255);
      enc->symbol(d0, u8_fold(diff_l));
    }
    if (sym & 2) {
      diff_h = (int)(c[0] >> 8) - (int)(last[0] >> 8);
      enc->symbol(d1, u8_fold(diff_h));
    }
    if (sym & 64) {
      if (sym & 4) {
        corr = (int)(c[1] & 255) - u8_clamp(diff_l + (int)(last[1] & 255));
        enc->symbol(d2, u8_fold(corr));
      }
      if (sym & 16) {
        diff_l = (diff_l + (int)(c[1] & 255) - (int)(last[1] & 255)) / 2;
        corr = (int)(c[2] & 255) - u8_clamp(diff_l + (int)(last[2] & 255));
        enc->symbol(d4, u8_fold(corr));
      }
      if (sym & 8) {
        corr = (int)(c[1] >> 8) - u8_clamp(diff_h + (int)(last[1] >> 8));
        enc->symbol(d3, u8_fold(corr));
      }
      if (sym & 32) {
        diff_h = (diff_h + (int)(c[1] >> 8) - (int)(last[1] >> 8)) / 2;
        corr = (int)(c[2] >> 8) - u8_clamp(diff_h + (int)(last[2] >> 8));
        enc->symbol(d5, u8_fold(corr));
      }
    }
    memcpy(last, c, 6);
  }
};

struct ByteV2Writer : ItemWriter {
  ArithmeticEncoder* enc;
  int number;
  std::vector<uint8_t> last;
  std::vector<std::unique_ptr<SymbolModel>> m_byte;
  ByteV2Writer(ArithmeticEncoder* e, int n) : enc(e), number(n), last((size_t)n) {
    for (int i = 0; i < n; ++i) m_byte.emplace_back(new SymbolModel(256));
  }
  void init(const uint8_t* first) override {
    for (auto& m : m_byte) m->init();
    memcpy(last.data(), first, (size_t)number);
  }
  void write(const uint8_t* item) override {
    for (int i = 0; i < number; ++i) enc->symbol(*m_byte[(size_t)i], u8_fold((int)item[i] - (int)last[(size_t)i]));
    memcpy(last.data(), item, (size_t)number);
  }
};


// ================================================================================================================
// LAS 1.4: "layered chunked" compression (compressor 3) of POINT14 v3 (+ BYTE14 v3 extra bytes) — the layout of the
// reference's tree_wind_condition.laz and als_multichannel_leg000_points.laz.  A chunk is: first point raw, u32 point
// count, the byte size of every layer, then the layers, each an arithmetic-coded stream of its own (channel / returns /
// XY, Z, classification, flags, intensity, scan angle, user data, point source, GPS time; one layer per extra byte).
// Up to four contexts (one per scanner channel) keep their own models and predictors.
// ================================================================================================================
// XY context of a return: single (0), first / last of two (1 / 2), first / intermediate / last of three or more (3 / 4 / 5).
// Pinned for every valid pair 1 <= r <= n by the reference's fixtures (decoded bounding boxes and by-return counts equal the
// headers, every layer's decoder ends on the last byte of its stream).  Points with r = 0 or r > n use the neighbouring class
// (not pinned by any fixture: such pairs only select which median predicts X and Y, they cannot desynchronise a stream).
inline uint32_t return_map6(uint32_t n, uint32_t r) {
  if (n <= 1) return 0;
  if (n == 2) return r <= 1 ? 1 : 2;
  return r <= 1 ? 3 : (r >= n ? 5 : 4);
}
inline uint32_t return_level8(uint32_t n, uint32_t r) {
  const uint32_t d = n > r ? n - r : r - n;
  return d < 7 ? d : 7;
}
constexpr int kGpsMulti = 500, kGpsMultiMinus = -10, kGpsMultiCodeFull = kGpsMulti - kGpsMultiMinus + 1, kGpsMultiTotal = kGpsMulti - kGpsMultiMinus + 5;

struct P14 {  // LAS point format 6 core, unpacked
  int32_t X, Y, Z;
  uint16_t intensity;
  uint8_t return_number, number_of_returns, class_flags, scanner_channel, scan_dir, edge, classification, user_data;
  int16_t scan_angle;
  uint16_t psid;
  double gps_time;
  bool gps_time_change;
  void unpack(const uint8_t* p) {
    X = rd_i32(p); Y = rd_i32(p + 4); Z = rd_i32(p + 8);
    intensity = rd_u16(p + 12);
    return_number = p[14] & 15; number_of_returns = p[14] >> 4;
    class_flags = p[15] & 15; scanner_channel = (p[15] >> 4) & 3; scan_dir = (p[15] >> 6) & 1; edge = p[15] >> 7;
    classification = p[16]; user_data = p[17];
    memcpy(&scan_angle, p + 18, 2);
    psid = rd_u16(p + 20);
    memcpy(&gps_time, p + 22, 8);
    gps_time_change = false;
  }
  void pack(uint8_t* p) const {
    wr_i32(p, X); wr_i32(p + 4, Y); wr_i32(p + 8, Z);
    wr_u16(p + 12, intensity);
    p[14] = (uint8_t)(return_number | (number_of_returns << 4));
    p[15] = (uint8_t)(class_flags | (scanner_channel << 4) | (scan_dir << 6) | (edge << 7));
    p[16] = classification; p[17] = user_data;
    memcpy(p + 18, &scan_angle, 2);
    wr_u16(p + 20, psid);
    memcpy(p + 22, &gps_time, 8);
  }
};

enum { kLayerXY = 0, kLayerZ, kLayerClass, kLayerFlags, kLayerIntensity, kLayerScanAngle, kLayerUserData, kLayerSource, kLayerGps, kLayers14 };

struct Point14V3 {
  struct Ctx {
    bool unused = true;
    P14 last;
    uint16_t last_intensity[8];
    Median5 mx[12], my[12];
    int32_t last_Z[8];
    std::unique_ptr<SymbolModel> changed_values[8], scanner_channel, number_of_returns[16], return_number_gps_same, return_number[16];
    std::unique_ptr<SymbolModel> classification[64], flags[64], user_data[64], gps_multi, gps_0diff;
    std::unique_ptr<IntegerDecompressor> ic_dx, ic_dy, ic_z, ic_intensity, ic_scan_angle, ic_source, ic_gps;
    uint32_t gps_last = 0, gps_next = 0;
    uint64_t last_gps[4];
    int32_t last_gps_diff[4], multi_extreme[4];
  };
  ByteReader br[kLayers14];
  ArithmeticDecoder dec[kLayers14];
  bool changed[kLayers14];
  Ctx ctx[4];
  uint32_t current = 0;

  static SymbolModel& lazy(std::unique_ptr<SymbolModel>& m, uint32_t n) {
    if (!m) m.reset(new SymbolModel(n));
    return *m;
  }
  void create_and_init(uint32_t c, const P14& item) {
    Ctx& x = ctx[c];
    for (int i = 0; i < 8; ++i) lazy(x.changed_values[i], 128).init();
    lazy(x.scanner_channel, 3).init();
    for (int i = 0; i < 16; ++i) {
      if (x.number_of_returns[i]) x.number_of_returns[i]->init();
      if (x.return_number[i]) x.return_number[i]->init();
    }
    lazy(x.return_number_gps_same, 13).init();
    if (!x.ic_dx) {
      x.ic_dx.reset(new IntegerDecompressor(&dec[kLayerXY], 32, 2));
      x.ic_dy.reset(new IntegerDecompressor(&dec[kLayerXY], 32, 22));
      x.ic_z.reset(new IntegerDecompressor(&dec[kLayerZ], 32, 20));
      x.ic_intensity.reset(new IntegerDecompressor(&dec[kLayerIntensity], 16, 4));
      x.ic_scan_angle.reset(new IntegerDecompressor(&dec[kLayerScanAngle], 16, 2));
      x.ic_source.reset(new IntegerDecompressor(&dec[kLayerSource], 16, 1));
      x.ic_gps.reset(new IntegerDecompressor(&dec[kLayerGps], 32, 9));
    }
    x.ic_dx->init(); x.ic_dy->init(); x.ic_z->init(); x.ic_intensity->init(); x.ic_scan_angle->init(); x.ic_source->init(); x.ic_gps->init();
    for (int i = 0; i < 12; ++i) { x.mx[i].init(); x.my[i].init(); }
    for (int i = 0; i < 8; ++i) { x.last_Z[i] = item.Z; x.last_intensity[i] = item.intensity; }
    for (int i = 0; i < 64; ++i) {
      if (x.classification[i]) x.classification[i]->init();
      if (x.flags[i]) x.flags[i]->init();
      if (x.user_data[i]) x.user_data[i]->init();
    }
    lazy(x.gps_multi, kGpsMultiTotal).init();
    lazy(x.gps_0diff, 5).init();
    x.gps_last = x.gps_next = 0;
    for (int i = 0; i < 4; ++i) { x.last_gps[i] = 0; x.last_gps_diff[i] = 0; x.multi_extreme[i] = 0; }
    memcpy(&x.last_gps[0], &item.gps_time, 8);
    x.last = item;
    x.last.gps_time_change = false;
    x.unused = false;
  }
  // layer sizes follow the chunk's point count; returns the number of bytes of all layers
  int64_t begin_chunk(const uint8_t* first_raw, const uint8_t* sizes, const uint8_t* layers, const uint8_t* limit) {
    uint32_t nbytes[kLayers14];
    memcpy(nbytes, sizes, sizeof(nbytes));
    const uint8_t* p = layers;
    for (int l = 0; l < kLayers14; ++l) {
      br[l] = ByteReader{p, p + nbytes[l] <= limit ? p + nbytes[l] : limit};
      changed[l] = nbytes[l] > 0;
      if (changed[l]) dec[l].init(&br[l]);
      p += nbytes[l];
    }
    for (auto& c : ctx) c.unused = true;
    P14 item;
    item.unpack(first_raw);
    current = item.scanner_channel;
    create_and_init(current, item);
    return p - layers;
  }
  void read_gps(Ctx& x) {
    ArithmeticDecoder& d = dec[kLayerGps];
    if (x.last_gps_diff[x.gps_last] == 0) {
      const int multi = (int)d.symbol(*x.gps_0diff);
      if (multi == 0) {  // the difference fits 32 bits
        x.last_gps_diff[x.gps_last] = x.ic_gps->decompress(0, 0);
        x.last_gps[x.gps_last] += (uint64_t)(int64_t)x.last_gps_diff[x.gps_last];
        x.multi_extreme[x.gps_last] = 0;
      } else if (multi == 1) {  // a huge difference: start another sequence
        x.gps_next = (x.gps_next + 1) & 3;
        uint64_t v = (uint64_t)(uint32_t)x.ic_gps->decompress((int32_t)(x.last_gps[x.gps_last] >> 32), 8);
        v = (v << 32) | d.read_bits(32);
        x.last_gps[x.gps_next] = v;
        x.gps_last = x.gps_next;
        x.last_gps_diff[x.gps_last] = 0;
        x.multi_extreme[x.gps_last] = 0;
      } else {  // switch to another sequence
        x.gps_last = (x.gps_last + (uint32_t)multi - 1) & 3;
        read_gps(x);
      }
    } else {
      int multi = (int)d.symbol(*x.gps_multi);
      if (multi == 1) {
        x.last_gps[x.gps_last] += (uint64_t)(int64_t)x.ic_gps->decompress(x.last_gps_diff[x.gps_last], 1);
        x.multi_extreme[x.gps_last] = 0;
      } else if (multi < kGpsMultiCodeFull) {
        int32_t diff;
        const int32_t ld = x.last_gps_diff[x.gps_last];
        if (multi == 0) {
          diff = x.ic_gps->decompress(0, 7);
          if (++x.multi_extreme[x.gps_last] > 3) { x.last_gps_diff[x.gps_last] = diff; x.multi_extreme[x.gps_last] = 0; }
        } else if (multi < kGpsMulti) {
          diff = x.ic_gps->decompress((int32_t)((uint32_t)multi * (uint32_t)ld), multi < 10 ? 2 : 3);
        } else if (multi == kGpsMulti) {
          diff = x.ic_gps->decompress((int32_t)((uint32_t)kGpsMulti * (uint32_t)ld), 4);
          if (++x.multi_extreme[x.gps_last] > 3) { x.last_gps_diff[x.gps_last] = diff; x.multi_extreme[x.gps_last] = 0; }
        } else {
          multi = kGpsMulti - multi;
          if (multi > kGpsMultiMinus) {
            diff = x.ic_gps->decompress((int32_t)((uint32_t)multi * (uint32_t)ld), 5);
          } else {
            diff = x.ic_gps->decompress((int32_t)((uint32_t)kGpsMultiMinus * (uint32_t)ld), 6);
            if (++x.multi_extreme[x.gps_last] > 3) { x.last_gps_diff[x.gps_last] = diff; x.multi_extreme[x.gps_last] = 0; }
          }
        }
        x.last_gps[x.gps_last] += (uint64_t)(int64_t)diff;
      } else if (multi == kGpsMultiCodeFull) {
        x.gps_next = (x.gps_next + 1) & 3;
        uint64_t v = (uint64_t)(uint32_t)x.ic_gps->decompress((int32_t)(x.last_gps[x.gps_last] >> 32), 8);
        v = (v << 32) | d.read_bits(32);
        x.last_gps[x.gps_next] = v;
        x.gps_last = x.gps_next;
        x.last_gps_diff[x.gps_last] = 0;
        x.multi_extreme[x.gps_last] = 0;
      } else {
        x.gps_last = (x.gps_last + (uint32_t)multi - (uint32_t)kGpsMultiCodeFull) & 3;
        read_gps(x);
      }
    }
  }
  // decodes the next point into out (30 bytes); returns the context (scanner channel) for the extra-byte item
  uint32_t read(uint8_t* out) {
    Ctx* x = &ctx[current];
    int lpr = (x->last.return_number == 1 ? 1 : 0) + (x->last.return_number >= x->last.number_of_returns ? 2 : 0) + (x->last.gps_time_change ? 4 : 0);
    const uint32_t cv = dec[kLayerXY].symbol(*x->changed_values[lpr]);
    if (cv & 64) {
      const uint32_t diff = dec[kLayerXY].symbol(*x->scanner_channel);
      const uint32_t sc = (current + diff + 1) % 4;
      if (ctx[sc].unused) create_and_init(sc, x->last);
      current = sc;
      x = &ctx[current];
      x->last.scanner_channel = (uint8_t)sc;
    }
    const bool source_change = cv & 32, gps_change = cv & 16, angle_change = cv & 8;
    const uint32_t last_n = x->last.number_of_returns, last_r = x->last.return_number;
    uint32_t n = last_n, r;
    if (cv & 4) {
      n = dec[kLayerXY].symbol(lazy(x->number_of_returns[last_n], 16));
      x->last.number_of_returns = (uint8_t)n;
    }
    switch (cv & 3) {
      case 0: r = last_r; break;
      case 1: r = (last_r + 1) % 16; break;
      case 2: r = (last_r + 15) % 16; break;
      default:
        if (gps_change) r = dec[kLayerXY].symbol(lazy(x->return_number[last_r], 16));
        else r = (last_r + dec[kLayerXY].symbol(*x->return_number_gps_same) + 2) % 16;
    }
    x->last.return_number = (uint8_t)r;
    const uint32_t m = return_map6(n, r);
    const uint32_t l = return_level8(n, r);
    const int cpr = (r == 1 ? 2 : 0) + (r >= n ? 1 : 0);
    const uint32_t mi = (m << 1) | (gps_change ? 1u : 0u);
    int32_t diff = x->ic_dx->decompress(x->mx[mi].get(), n == 1);
    x->last.X = (int32_t)((uint32_t)x->last.X + (uint32_t)diff);
    x->mx[mi].add(diff);
    uint32_t kb = x->ic_dx->k;
    diff = x->ic_dy->decompress(x->my[mi].get(), (n == 1) + (kb < 20 ? (kb & ~1u) : 20));
    x->last.Y = (int32_t)((uint32_t)x->last.Y + (uint32_t)diff);
    x->my[mi].add(diff);
    if (changed[kLayerZ]) {
      kb = (x->ic_dx->k + x->ic_dy->k) / 2;
      x->last.Z = x->ic_z->decompress(x->last_Z[l], (n == 1) + (kb < 18 ? (kb & ~1u) : 18));
      x->last_Z[l] = x->last.Z;
    }
    if (changed[kLayerClass]) {
      const int ccc = ((x->last.classification & 0x1F) << 1) + (cpr == 3 ? 1 : 0);
      x->last.classification = (uint8_t)dec[kLayerClass].symbol(lazy(x->classification[ccc], 256));
    }
    if (changed[kLayerFlags]) {
      const uint32_t lf = ((uint32_t)x->last.edge << 5) | ((uint32_t)x->last.scan_dir << 4) | x->last.class_flags;
      const uint32_t f = dec[kLayerFlags].symbol(lazy(x->flags[lf], 64));
      x->last.edge = (f >> 5) & 1;
      x->last.scan_dir = (f >> 4) & 1;
      x->last.class_flags = f & 15;
    }
    if (changed[kLayerIntensity]) {
      const uint32_t ii = ((uint32_t)cpr << 1) | (gps_change ? 1u : 0u);
      const uint16_t v = (uint16_t)x->ic_intensity->decompress(x->last_intensity[ii], (uint32_t)cpr);
      x->last_intensity[ii] = v;
      x->last.intensity = v;
    }
    if (changed[kLayerScanAngle] && angle_change)
      x->last.scan_angle = (int16_t)x->ic_scan_angle->decompress((int32_t)(uint16_t)x->last.scan_angle, gps_change ? 1 : 0);
    if (changed[kLayerUserData]) x->last.user_data = (uint8_t)dec[kLayerUserData].symbol(lazy(x->user_data[x->last.user_data / 4], 256));
    if (changed[kLayerSource] && source_change) x->last.psid = (uint16_t)x->ic_source->decompress(x->last.psid, 0);
    if (changed[kLayerGps] && gps_change) {
      read_gps(*x);
      memcpy(&x->last.gps_time, &x->last_gps[x->gps_last], 8);
    }
    x->last.pack(out);
    x->last.gps_time_change = gps_change;
    return current;
  }
  bool overrun() const {
    for (int l = 0; l < kLayers14; ++l)
      if (changed[l] && br[l].overrun) return true;
    return false;
  }
};

struct Byte14V3 {
  int number;
  struct Ctx {
    bool unused = true;
    std::vector<uint8_t> last;
    std::vector<std::unique_ptr<SymbolModel>> m;
  };
  std::vector<ByteReader> br;
  std::vector<ArithmeticDecoder> dec;
  std::vector<char> changed;
  Ctx ctx[4];
  uint32_t current = 0;
  explicit Byte14V3(int n) : number(n), br((size_t)n), dec((size_t)n), changed((size_t)n) {}
  void create_and_init(uint32_t c, const uint8_t* item) {
    Ctx& x = ctx[c];
    if (x.m.empty())
      for (int i = 0; i < number; ++i) x.m.emplace_back(new SymbolModel(256));
    for (auto& m : x.m) m->init();
    x.last.assign(item, item + number);
    x.unused = false;
  }
  int64_t begin_chunk(const uint8_t* first_raw, uint32_t context, const uint8_t* sizes, const uint8_t* layers, const uint8_t* limit) {
    const uint8_t* p = layers;
    for (int i = 0; i < number; ++i) {
      uint32_t nb;
      memcpy(&nb, sizes + 4 * i, 4);
      br[(size_t)i] = ByteReader{p, p + nb <= limit ? p + nb : limit};
      changed[(size_t)i] = nb > 0;
      if (nb) dec[(size_t)i].init(&br[(size_t)i]);
      p += nb;
    }
    for (auto& c : ctx) c.unused = true;
    current = context;
    create_and_init(current, first_raw);
    return p - layers;
  }
  void read(uint8_t* out, uint32_t context) {
    if (context != current) {
      if (ctx[context].unused) create_and_init(context, ctx[current].last.data());
      current = context;
    }
    Ctx& x = ctx[current];
    for (int i = 0; i < number; ++i) {
      if (changed[(size_t)i]) x.last[(size_t)i] = u8_fold((int)x.last[(size_t)i] + (int)dec[(size_t)i].symbol(*x.m[(size_t)i]));
      out[i] = x.last[(size_t)i];
    }
  }
  bool overrun() const {
    for (int i = 0; i < number; ++i)
      if (changed[(size_t)i] && br[(size_t)i].overrun) return true;
    return false;
  }
};


// RGB14 v3 (LAS 1.4 point formats 7 / 8): the colour coder of RGB12 v2 in a layer of its own, one context per scanner channel.
// No fixture of the reference holds this item: round-trip tested, and guarded like every layer by the end-of-stream check.
struct Rgb14V3 {
  ByteReader br;
  ArithmeticDecoder dec;
  bool changed = false;
  std::unique_ptr<Rgb12V2> ctx[4];
  bool unused[4];
  uint32_t current = 0;
  void create_and_init(uint32_t c, const uint8_t* item) {
    if (!ctx[c]) ctx[c].reset(new Rgb12V2(&dec));
    ctx[c]->init(item);
    unused[c] = false;
  }
  int64_t begin_chunk(const uint8_t* first_raw, uint32_t context, const uint8_t* sizes, const uint8_t* layers, const uint8_t* limit) {
    uint32_t nb;
    memcpy(&nb, sizes, 4);
    br = ByteReader{layers, layers + nb <= limit ? layers + nb : limit};
    changed = nb > 0;
    if (changed) dec.init(&br);
    for (auto& u : unused) u = true;
    current = context;
    create_and_init(current, first_raw);
    return nb;
  }
  void read(uint8_t* out, uint32_t context) {
    if (context != current) {
      if (unused[context]) create_and_init(context, reinterpret_cast<const uint8_t*>(ctx[current]->last));
      current = context;
    }
    if (changed) ctx[current]->read(out);
    else memcpy(out, ctx[current]->last, 6);
  }
};
struct Rgb14V3Writer {
  std::vector<uint8_t> buf;
  ArithmeticEncoder enc;
  bool changed = false;
  std::unique_ptr<Rgb12V2Writer> ctx[4];
  bool unused[4];
  uint32_t current = 0;
  void create_and_init(uint32_t c, const uint8_t* item) {
    if (!ctx[c]) ctx[c].reset(new Rgb12V2Writer(&enc));
    ctx[c]->init(item);
    unused[c] = false;
  }
  void begin_chunk(const uint8_t* first_raw, uint32_t context) {
    buf.clear();
    enc.init(&buf);
    changed = false;
    for (auto& u : unused) u = true;
    current = context;
    create_and_init(current, first_raw);
  }
  void write(const uint8_t* item, uint32_t context) {
    if (context != current) {
      if (unused[context]) create_and_init(context, reinterpret_cast<const uint8_t*>(ctx[current]->last));
      current = context;
    }
    if (memcmp(ctx[current]->last, item, 6) != 0) changed = true;
    ctx[current]->write(item);
  }
  void end_chunk() { enc.done(); }
  uint32_t layer_bytes() const { return changed ? (uint32_t)buf.size() : 0u; }
};

// ---- encoder side of the layered items: the mirror of Point14V3 / Byte14V3 ------------------------------------------
struct Point14V3Writer {
  struct Ctx {
    bool unused = true;
    P14 last;
    uint16_t last_intensity[8];
    Median5 mx[12], my[12];
    int32_t last_Z[8];
    std::unique_ptr<SymbolModel> changed_values[8], scanner_channel, number_of_returns[16], return_number_gps_same, return_number[16];
    std::unique_ptr<SymbolModel> classification[64], flags[64], user_data[64], gps_multi, gps_0diff;
    std::unique_ptr<IntegerCompressor> ic_dx, ic_dy, ic_z, ic_intensity, ic_scan_angle, ic_source, ic_gps;
    uint32_t gps_last = 0, gps_next = 0;
    uint64_t last_gps[4];
    int32_t last_gps_diff[4], multi_extreme[4];
  };
  std::vector<uint8_t> buf[kLayers14];
  ArithmeticEncoder enc[kLayers14];
  bool changed[kLayers14];
  Ctx ctx[4];
  uint32_t current = 0;

  static SymbolModel& lazy(std::unique_ptr<SymbolModel>& m, uint32_t n) { return Point14V3::lazy(m, n); }
  void create_and_init(uint32_t c, const P14& item) {
    Ctx& x = ctx[c];
    for (int i = 0; i < 8; ++i) lazy(x.changed_values[i], 128).init();
    lazy(x.scanner_channel, 3).init();
    for (int i = 0; i < 16; ++i) {
      if (x.number_of_returns[i]) x.number_of_returns[i]->init();
      if (x.return_number[i]) x.return_number[i]->init();
    }
    lazy(x.return_number_gps_same, 13).init();
    if (!x.ic_dx) {
      x.ic_dx.reset(new IntegerCompressor(&enc[kLayerXY], 32, 2));
      x.ic_dy.reset(new IntegerCompressor(&enc[kLayerXY], 32, 22));
      x.ic_z.reset(new IntegerCompressor(&enc[kLayerZ], 32, 20));
      x.ic_intensity.reset(new IntegerCompressor(&enc[kLayerIntensity], 16, 4));
      x.ic_scan_angle.reset(new IntegerCompressor(&enc[kLayerScanAngle], 16, 2));
      x.ic_source.reset(new IntegerCompressor(&enc[kLayerSource], 16, 1));
      x.ic_gps.reset(new IntegerCompressor(&enc[kLayerGps], 32, 9));
    }
    x.ic_dx->init(); x.ic_dy->init(); x.ic_z->init(); x.ic_intensity->init(); x.ic_scan_angle->init(); x.ic_source->init(); x.ic_gps->init();
    for (int i = 0; i < 12; ++i) { x.mx[i].init(); x.my[i].init(); }
    for (int i = 0; i < 8; ++i) { x.last_Z[i] = item.Z; x.last_intensity[i] = item.intensity; }
    for (int i = 0; i < 64; ++i) {
      if (x.classification[i]) x.classification[i]->init();
      if (x.flags[i]) x.flags[i]->init();
      if (x.user_data[i]) x.user_data[i]->init();
    }
    lazy(x.gps_multi, kGpsMultiTotal).init();
    lazy(x.gps_0diff, 5).init();
    x.gps_last = x.gps_next = 0;
    for (int i = 0; i < 4; ++i) { x.last_gps[i] = 0; x.last_gps_diff[i] = 0; x.multi_extreme[i] = 0; }
    memcpy(&x.last_gps[0], &item.gps_time, 8);
    x.last = item;
    x.last.gps_time_change = false;
    x.unused = false;
  }
  uint32_t begin_chunk(const uint8_t* first_raw) {
    for (int l = 0; l < kLayers14; ++l) {
      buf[l].clear();
      enc[l].init(&buf[l]);
      changed[l] = false;
    }
    for (auto& c : ctx) c.unused = true;
    P14 item;
    item.unpack(first_raw);
    current = item.scanner_channel;
    create_and_init(current, item);
    return current;
  }
  static int32_t quantize(float f) { return f >= 0 ? (int32_t)(f + 0.5f) : (int32_t)(f - 0.5f); }
  void write_gps(Ctx& x, uint64_t gps) {
    ArithmeticEncoder& e = enc[kLayerGps];
    const int64_t d64 = (int64_t)(gps - x.last_gps[x.gps_last]);
    const int32_t d32 = (int32_t)d64;
    if (x.last_gps_diff[x.gps_last] == 0) {
      if (d64 == (int64_t)d32) {
        e.symbol(*x.gps_0diff, 0);
        x.ic_gps->compress(0, d32, 0);
        x.last_gps_diff[x.gps_last] = d32;
        x.multi_extreme[x.gps_last] = 0;
      } else {
        for (uint32_t i = 1; i < 4; ++i) {
          const int64_t o64 = (int64_t)(gps - x.last_gps[(x.gps_last + i) & 3]);
          if (o64 == (int64_t)(int32_t)o64) {
            e.symbol(*x.gps_0diff, i + 1);
            x.gps_last = (x.gps_last + i) & 3;
            write_gps(x, gps);
            return;
          }
        }
        e.symbol(*x.gps_0diff, 1);
        x.ic_gps->compress((int32_t)(x.last_gps[x.gps_last] >> 32), (int32_t)(gps >> 32), 8);
        e.write_bits(32, (uint32_t)gps);
        x.gps_next = (x.gps_next + 1) & 3;
        x.gps_last = x.gps_next;
        x.last_gps_diff[x.gps_last] = 0;
        x.multi_extreme[x.gps_last] = 0;
      }
      x.last_gps[x.gps_last] = gps;
    } else {
      if (d64 == (int64_t)d32) {
        const int32_t ld = x.last_gps_diff[x.gps_last];
        const int32_t multi = quantize((float)d32 / (float)ld);
        if (multi == 1) {
          e.symbol(*x.gps_multi, 1);
          x.ic_gps->compress(ld, d32, 1);
          x.multi_extreme[x.gps_last] = 0;
        } else if (multi > 0) {
          if (multi < kGpsMulti) {
            e.symbol(*x.gps_multi, (uint32_t)multi);
            x.ic_gps->compress((int32_t)((uint32_t)multi * (uint32_t)ld), d32, multi < 10 ? 2 : 3);
          } else {
            e.symbol(*x.gps_multi, kGpsMulti);
            x.ic_gps->compress((int32_t)((uint32_t)kGpsMulti * (uint32_t)ld), d32, 4);
            if (++x.multi_extreme[x.gps_last] > 3) { x.last_gps_diff[x.gps_last] = d32; x.multi_extreme[x.gps_last] = 0; }
          }
        } else if (multi < 0) {
          if (multi > kGpsMultiMinus) {
            e.symbol(*x.gps_multi, (uint32_t)(kGpsMulti - multi));
            x.ic_gps->compress((int32_t)((uint32_t)multi * (uint32_t)ld), d32, 5);
          } else {
            e.symbol(*x.gps_multi, (uint32_t)(kGpsMulti - kGpsMultiMinus));
            x.ic_gps->compress((int32_t)((uint32_t)kGpsMultiMinus * (uint32_t)ld), d32, 6);
            if (++x.multi_extreme[x.gps_last] > 3) { x.last_gps_diff[x.gps_last] = d32; x.multi_extreme[x.gps_last] = 0; }
          }
        } else {
          e.symbol(*x.gps_multi, 0);
          x.ic_gps->compress(0, d32, 7);
          if (++x.multi_extreme[x.gps_last] > 3) { x.last_gps_diff[x.gps_last] = d32; x.multi_extreme[x.gps_last] = 0; }
        }
      } else {
        for (uint32_t i = 1; i < 4; ++i) {
          const int64_t o64 = (int64_t)(gps - x.last_gps[(x.gps_last + i) & 3]);
          if (o64 == (int64_t)(int32_t)o64) {
            e.symbol(*x.gps_multi, (uint32_t)kGpsMultiCodeFull + i);
            x.gps_last = (x.gps_last + i) & 3;
            write_gps(x, gps);
            return;
          }
        }
        e.symbol(*x.gps_multi, kGpsMultiCodeFull);
        x.ic_gps->compress((int32_t)(x.last_gps[x.gps_last] >> 32), (int32_t)(gps >> 32), 8);
        e.write_bits(32, (uint32_t)gps);
        x.gps_next = (x.gps_next + 1) & 3;
        x.gps_last = x.gps_next;
        x.last_gps_diff[x.gps_last] = 0;
        x.multi_extreme[x.gps_last] = 0;
      }
      x.last_gps[x.gps_last] = gps;
    }
  }
  uint32_t write(const uint8_t* raw) {
    P14 item;
    item.unpack(raw);
    Ctx* x = &ctx[current];
    const int lpr = (x->last.return_number == 1 ? 1 : 0) + (x->last.return_number >= x->last.number_of_returns ? 2 : 0) + (x->last.gps_time_change ? 4 : 0);
    const uint32_t sc = item.scanner_channel;
    const P14* ref = &x->last;  // what the point is compared with: the last point of ITS channel when that channel is alive
    if (sc != current && !ctx[sc].unused) ref = &ctx[sc].last;
    uint64_t gps_bits, ref_bits;
    memcpy(&gps_bits, &item.gps_time, 8);
    memcpy(&ref_bits, &ref->gps_time, 8);
    const bool source_change = item.psid != ref->psid, gps_change = gps_bits != ref_bits, angle_change = item.scan_angle != ref->scan_angle;
    const uint32_t last_n = ref->number_of_returns, last_r = ref->return_number, n = item.number_of_returns, r = item.return_number;
    uint32_t cv = ((uint32_t)(sc != current) << 6) | ((uint32_t)source_change << 5) | ((uint32_t)gps_change << 4) | ((uint32_t)angle_change << 3) |
                  ((uint32_t)(n != last_n) << 2);
    if (r != last_r) cv |= r == (last_r + 1) % 16 ? 1u : (r == (last_r + 15) % 16 ? 2u : 3u);
    enc[kLayerXY].symbol(*x->changed_values[lpr], cv);
    if (cv & 64) {
      const int diff = (int)sc - (int)current;
      enc[kLayerXY].symbol(*x->scanner_channel, (uint32_t)(diff > 0 ? diff - 1 : diff + 4 - 1));
      if (ctx[sc].unused) create_and_init(sc, x->last);
      current = sc;
      x = &ctx[current];
    }
    if (cv & 4) enc[kLayerXY].symbol(lazy(x->number_of_returns[last_n], 16), n);
    if ((cv & 3) == 3) {
      if (gps_change) {
        enc[kLayerXY].symbol(lazy(x->return_number[last_r], 16), r);
      } else {
        const int diff = (int)r - (int)last_r;
        enc[kLayerXY].symbol(*x->return_number_gps_same, (uint32_t)(diff > 1 ? diff - 2 : diff + 16 - 2));
      }
    }
    const uint32_t m = return_map6(n, r), l = return_level8(n, r);
    const int cpr = (r == 1 ? 2 : 0) + (r >= n ? 1 : 0);
    const uint32_t mi = (m << 1) | (gps_change ? 1u : 0u);
    int32_t diff = (int32_t)((uint32_t)item.X - (uint32_t)x->last.X);
    x->ic_dx->compress(x->mx[mi].get(), diff, n == 1);
    x->mx[mi].add(diff);
    uint32_t kb = x->ic_dx->k;
    diff = (int32_t)((uint32_t)item.Y - (uint32_t)x->last.Y);
    x->ic_dy->compress(x->my[mi].get(), diff, (n == 1) + (kb < 20 ? (kb & ~1u) : 20));
    x->my[mi].add(diff);
    kb = (x->ic_dx->k + x->ic_dy->k) / 2;
    x->ic_z->compress(x->last_Z[l], item.Z, (n == 1) + (kb < 18 ? (kb & ~1u) : 18));
    if (x->last_Z[l] != item.Z) changed[kLayerZ] = true;
    x->last_Z[l] = item.Z;
    const int ccc = ((x->last.classification & 0x1F) << 1) + (cpr == 3 ? 1 : 0);
    enc[kLayerClass].symbol(lazy(x->classification[ccc], 256), item.classification);
    if (x->last.classification != item.classification) changed[kLayerClass] = true;
    const uint32_t lf = ((uint32_t)x->last.edge << 5) | ((uint32_t)x->last.scan_dir << 4) | x->last.class_flags;
    const uint32_t f = ((uint32_t)item.edge << 5) | ((uint32_t)item.scan_dir << 4) | item.class_flags;
    enc[kLayerFlags].symbol(lazy(x->flags[lf], 64), f);
    if (lf != f) changed[kLayerFlags] = true;
    const uint32_t ii = ((uint32_t)cpr << 1) | (gps_change ? 1u : 0u);
    x->ic_intensity->compress(x->last_intensity[ii], item.intensity, (uint32_t)cpr);
    if (x->last_intensity[ii] != item.intensity) changed[kLayerIntensity] = true;
    x->last_intensity[ii] = item.intensity;
    if (angle_change) {
      x->ic_scan_angle->compress((int32_t)(uint16_t)x->last.scan_angle, (int32_t)(uint16_t)item.scan_angle, gps_change ? 1 : 0);
      changed[kLayerScanAngle] = true;
    }
    enc[kLayerUserData].symbol(lazy(x->user_data[x->last.user_data / 4], 256), item.user_data);
    if (x->last.user_data != item.user_data) changed[kLayerUserData] = true;
    if (source_change) {
      x->ic_source->compress(x->last.psid, item.psid, 0);
      changed[kLayerSource] = true;
    }
    if (gps_change) {
      write_gps(*x, gps_bits);
      changed[kLayerGps] = true;
    }
    x->last = item;
    x->last.gps_time_change = gps_change;
    return current;
  }
  void end_chunk() {
    for (int l = 0; l < kLayers14; ++l) enc[l].done();
  }
  uint32_t layer_bytes(int l) const { return (l == kLayerXY || changed[l]) ? (uint32_t)buf[l].size() : 0u; }
};

struct Byte14V3Writer {
  int number;
  struct Ctx {
    bool unused = true;
    std::vector<uint8_t> last;
    std::vector<std::unique_ptr<SymbolModel>> m;
  };
  std::vector<std::vector<uint8_t>> buf;
  std::vector<ArithmeticEncoder> enc;
  std::vector<char> changed;
  Ctx ctx[4];
  uint32_t current = 0;
  explicit Byte14V3Writer(int n) : number(n), buf((size_t)n), enc((size_t)n), changed((size_t)n) {}
  void create_and_init(uint32_t c, const uint8_t* item) {
    Ctx& x = ctx[c];
    if (x.m.empty())
      for (int i = 0; i < number; ++i) x.m.emplace_back(new SymbolModel(256));
    for (auto& m : x.m) m->init();
    x.last.assign(item, item + number);
    x.unused = false;
  }
  void begin_chunk(const uint8_t* first_raw, uint32_t context) {
    for (int i = 0; i < number; ++i) {
      buf[(size_t)i].clear();
      enc[(size_t)i].init(&buf[(size_t)i]);
      changed[(size_t)i] = 0;
    }
    for (auto& c : ctx) c.unused = true;
    current = context;
    create_and_init(current, first_raw);
  }
  void write(const uint8_t* item, uint32_t context) {
    if (context != current) {
      if (ctx[context].unused) create_and_init(context, ctx[current].last.data());
      current = context;
    }
    Ctx& x = ctx[current];
    for (int i = 0; i < number; ++i) {
      const int diff = (int)item[i] - (int)x.last[(size_t)i];
      enc[(size_t)i].symbol(*x.m[(size_t)i], u8_fold(diff));
      if (diff) {
        changed[(size_t)i] = 1;
        x.last[(size_t)i] = item[i];
      }
    }
  }
  void end_chunk() {
    for (int i = 0; i < number; ++i) enc[(size_t)i].done();
  }
  uint32_t layer_bytes(int i) const { return changed[(size_t)i] ? (uint32_t)buf[(size_t)i].size() : 0u; }
};

void write_chunk_table(std::vector<uint8_t>& out, const std::vector<uint32_t>& chunk_bytes) {
  const uint32_t version = 0, count = (uint32_t)chunk_bytes.size();
  const uint8_t* v = reinterpret_cast<const uint8_t*>(&version);
  const uint8_t* n = reinterpret_cast<const uint8_t*>(&count);
  out.insert(out.end(), v, v + 4);
  out.insert(out.end(), n, n + 4);
  if (count == 0) return;
  ArithmeticEncoder enc;
  enc.init(&out);
  IntegerCompressor ic(&enc, 32, 2);
  int32_t prev = 0;
  for (uint32_t b : chunk_bytes) {
    ic.compress(prev, (int32_t)b, 1);
    prev = (int32_t)b;
  }
  enc.done();
}

int encode_layered(const uint8_t* records, int64_t npoints, int32_t point_size, uint32_t chunk_size, int rgb, int extra, std::vector<uint8_t>& out,
                   std::vector<uint32_t>& chunk_bytes) {
  struct State {
    std::unique_ptr<Point14V3Writer> p14;
    std::unique_ptr<Rgb14V3Writer> c14;
    std::unique_ptr<Byte14V3Writer> b14;
  };
  auto make = [&]() {
    std::unique_ptr<State> st(new State());
    st->p14.reset(new Point14V3Writer());
    if (rgb) st->c14.reset(new Rgb14V3Writer());
    if (extra) st->b14.reset(new Byte14V3Writer(extra));
    return st;
  };
  const int64_t nchunks = (npoints + chunk_size - 1) / chunk_size;
  std::vector<std::vector<uint8_t>> parts((size_t)nchunks);
  auto one_chunk = [&](int64_t c, State& st) -> int {
    Point14V3Writer* p14 = st.p14.get();
    Rgb14V3Writer* c14 = st.c14.get();
    Byte14V3Writer* b14 = st.b14.get();
    std::vector<uint8_t>& o = parts[(size_t)c];
    const int64_t first = c * (int64_t)chunk_size;
    const int64_t count = std::min<int64_t>(chunk_size, npoints - first);
    const uint8_t* rec = records + first * point_size;
    o.insert(o.end(), rec, rec + point_size);
    const uint32_t context = p14->begin_chunk(rec);
    if (c14) c14->begin_chunk(rec + 30, context);
    if (b14) b14->begin_chunk(rec + 30 + rgb, context);
    for (int64_t p = 1; p < count; ++p) {
      const uint8_t* r = rec + p * point_size;
      const uint32_t cx = p14->write(r);
      if (c14) c14->write(r + 30, cx);
      if (b14) b14->write(r + 30 + rgb, cx);
    }
    p14->end_chunk();
    if (c14) c14->end_chunk();
    if (b14) b14->end_chunk();
    auto put32 = [&](uint32_t v) {
      const uint8_t* q = reinterpret_cast<const uint8_t*>(&v);
      o.insert(o.end(), q, q + 4);
    };
    put32((uint32_t)count);
    for (int l = 0; l < kLayers14; ++l) put32(p14->layer_bytes(l));
    if (c14) put32(c14->layer_bytes());
    for (int i = 0; i < extra; ++i) put32(b14->layer_bytes(i));
    for (int l = 0; l < kLayers14; ++l)
      if (p14->layer_bytes(l)) o.insert(o.end(), p14->buf[l].begin(), p14->buf[l].end());
    if (c14 && c14->layer_bytes()) o.insert(o.end(), c14->buf.begin(), c14->buf.end());
    for (int i = 0; i < extra; ++i)
      if (b14->layer_bytes(i)) o.insert(o.end(), b14->buf[(size_t)i].begin(), b14->buf[(size_t)i].end());
    return VAPC_OK;
  };
  if (const int rc = parallel_chunks(nchunks, make, one_chunk)) return rc;
  for (auto& part : parts) {
    out.insert(out.end(), part.begin(), part.end());
    chunk_bytes.push_back((uint32_t)part.size());
    std::vector<uint8_t>().swap(part);
  }
  return VAPC_OK;
}

// chunk start positions from the chunk table (both compressors write the same table)
int read_chunk_table(const char* who, const uint8_t* file, int64_t file_bytes, int64_t offset_to_points, int64_t nchunks, std::vector<int64_t>& chunk_start) {
  int64_t table_pos;
  memcpy(&table_pos, file + offset_to_points, 8);
  if (table_pos == -1) memcpy(&table_pos, file + file_bytes - 8, 8);
  chunk_start.assign((size_t)nchunks + 1, 0);
  chunk_start[0] = offset_to_points + 8;
  if (table_pos < chunk_start[0] || table_pos + 8 > file_bytes) return vapc_set_error(VAPC_E_INVALID, "%s: chunk table missing", who);
  uint32_t version, count;
  memcpy(&version, file + table_pos, 4);
  memcpy(&count, file + table_pos + 4, 4);
  if (version != 0 || (int64_t)count < nchunks) return vapc_set_error(VAPC_E_INVALID, "%s: unexpected chunk table (version %u, %u chunks)", who, version, count);
  ByteReader br{file + table_pos + 8, file + file_bytes};
  ArithmeticDecoder dec;
  dec.init(&br);
  IntegerDecompressor ic(&dec, 32, 2);
  int32_t prev = 0;
  for (int64_t i = 0; i < nchunks; ++i) {
    prev = ic.decompress(prev, 1);
    chunk_start[(size_t)i + 1] = chunk_start[(size_t)i] + (uint32_t)prev;
  }
  if (br.overrun) return vapc_set_error(VAPC_E_INVALID, "%s: chunk table truncated", who);
  return VAPC_OK;
}

// items of a layered file: POINT14 v3 [+ RGB14 v3] [+ BYTE14 v3]
int layered_items(const char* who, const uint16_t* items, int nitems, int point_size, int* rgb, int* extra) {
  *rgb = *extra = 0;
  bool ok = nitems >= 1 && nitems <= 3 && items[0] == 10 && items[1] == 30 && items[2] == 3;
  int bad = 0;
  for (int i = 1; ok && i < nitems; ++i) {
    const int type = items[3 * i], size = items[3 * i + 1], version = items[3 * i + 2];
    if (type == 11 && size == 6 && version == 3 && i == 1) *rgb = 6;
    else if (type == 14 && size > 0 && version == 3 && i == nitems - 1) *extra = size;
    else { ok = false; bad = i; }
  }
  if (!ok)
    return vapc_set_error(VAPC_E_INVALID, "%s: layered LASzip item type %d (size %d, version %d) is not supported (POINT14 v3, RGB14 v3 and BYTE14 v3 are)", who,
                          items[3 * bad], items[3 * bad + 1], items[3 * bad + 2]);
  if (30 + *rgb + *extra != point_size) return vapc_set_error(VAPC_E_INVALID, "%s: item sizes (%d) do not add up to the point size (%d)", who, 30 + *rgb + *extra, point_size);
  return VAPC_OK;
}

int decode_layered(const uint8_t* file, int64_t file_bytes, int64_t offset_to_points, int64_t npoints, int32_t point_size, uint32_t chunk_size,
                   const uint16_t* items, int32_t nitems, uint8_t* out) {
  const char* who = "vapc_laz_decode";
  int extra = 0, rgb = 0;
  if (const int rc = layered_items(who, items, nitems, point_size, &rgb, &extra)) return rc;
  if (npoints == 0) return VAPC_OK;
  if (chunk_size == 0 || chunk_size == 0xFFFFFFFFu) return vapc_set_error(VAPC_E_INVALID, "%s: variable chunk sizes are not supported", who);
  const int64_t nchunks = (npoints + chunk_size - 1) / chunk_size;
  std::vector<int64_t> chunk_start;
  if (const int rc = read_chunk_table(who, file, file_bytes, offset_to_points, nchunks, chunk_start)) return rc;
  struct State {
    std::unique_ptr<Point14V3> p14;
    std::unique_ptr<Rgb14V3> c14;
    std::unique_ptr<Byte14V3> b14;
  };
  auto make = [&]() {
    std::unique_ptr<State> st(new State());
    st->p14.reset(new Point14V3());
    if (rgb) st->c14.reset(new Rgb14V3());
    if (extra) st->b14.reset(new Byte14V3(extra));
    return st;
  };
  auto one_chunk = [&](int64_t c, State& st) -> int {
    Point14V3* p14 = st.p14.get();
    Rgb14V3* c14 = st.c14.get();
    Byte14V3* b14 = st.b14.get();
    const int64_t first = c * (int64_t)chunk_size;
    const int64_t count = std::min<int64_t>(chunk_size, npoints - first);
    const int64_t begin = chunk_start[(size_t)c], limit = chunk_start[(size_t)c + 1];
    const int64_t head = point_size + 4 + 4 * (kLayers14 + (rgb ? 1 : 0) + extra);
    if (limit > file_bytes || begin + head > limit) return vapc_set_error(VAPC_E_INVALID, "%s: chunk %lld outside the file", who, (long long)c);
    uint8_t* rec = out + first * point_size;
    memcpy(rec, file + begin, (size_t)point_size);  // the first point of a chunk is stored raw
    uint32_t stored;
    memcpy(&stored, file + begin + point_size, 4);
    if ((int64_t)stored != count) return vapc_set_error(VAPC_E_INVALID, "%s: chunk %lld holds %u points, expected %lld", who, (long long)c, stored, (long long)count);
    const uint8_t* sizes = file + begin + point_size + 4;
    const uint8_t* layers = file + begin + head;
    int64_t used = p14->begin_chunk(rec, sizes, layers, file + limit);
    if (c14) used += c14->begin_chunk(rec + 30, p14->current, sizes + 4 * kLayers14, layers + used, file + limit);
    if (b14) b14->begin_chunk(rec + 30 + rgb, p14->current, sizes + 4 * (kLayers14 + (rgb ? 1 : 0)), layers + used, file + limit);
    for (int64_t p = 1; p < count; ++p) {
      uint8_t* r = rec + p * point_size;
      const uint32_t context = p14->read(r);
      if (c14) c14->read(r + 30, context);
      if (b14) b14->read(r + 30 + rgb, context);
    }
    // a correct decode ends every layer on the last byte of its stream
    bool exact = !p14->overrun();
    for (int l = 0; l < kLayers14 && exact; ++l) exact = !p14->changed[l] || p14->br[l].p == p14->br[l].end;
    if (c14 && c14->changed) exact = exact && !c14->br.overrun && c14->br.p == c14->br.end;
    if (b14) {
      exact = exact && !b14->overrun();
      for (int i = 0; i < b14->number && exact; ++i) exact = !b14->changed[(size_t)i] || b14->br[(size_t)i].p == b14->br[(size_t)i].end;
    }
    if (!exact) return vapc_set_error(VAPC_E_INVALID, "%s: chunk %lld is truncated or corrupt (a layer's decoder did not end with its stream)", who, (long long)c);
    return VAPC_OK;
  };
  return parallel_chunks(nchunks, make, one_chunk);
}


// GPSTIME11 version 2 (LAS point formats 1 / 3): the GPS time coder of the layered format (pinned byte for byte by the LAS 1.4
// fixtures) with the two extra "unchanged" symbols version 2 needs, there being no separate change flag.  No fixture of the
// reference holds this item: round-trip tested, and guarded by the end-of-chunk check of the pointwise decoder.
constexpr int kGps2Unchanged = kGpsMulti - kGpsMultiMinus + 1, kGps2CodeFull = kGpsMulti - kGpsMultiMinus + 2, kGps2Total = kGpsMulti - kGpsMultiMinus + 6;
struct GpsState {
  uint32_t last = 0, next = 0;
  uint64_t time[4];
  int32_t diff[4], extreme[4];
  void init(const uint8_t* first) {
    last = next = 0;
    for (int i = 0; i < 4; ++i) { time[i] = 0; diff[i] = 0; extreme[i] = 0; }
    memcpy(&time[0], first, 8);
  }
  void bump(int32_t d) {
    if (++extreme[last] > 3) { diff[last] = d; extreme[last] = 0; }
  }
};
struct GpsTime11V2 : ItemReader {
  ArithmeticDecoder* dec;
  GpsState g;
  SymbolModel multi_model{(uint32_t)kGps2Total}, zero_model{6};
  IntegerDecompressor ic;
  explicit GpsTime11V2(ArithmeticDecoder* d) : dec(d), ic(d, 32, 9) {}
  void init(const uint8_t* first) override {
    g.init(first);
    multi_model.init();
    zero_model.init();
    ic.init();
  }
  void new_sequence() {
    g.next = (g.next + 1) & 3;
    uint64_t v = (uint64_t)(uint32_t)ic.decompress((int32_t)(g.time[g.last] >> 32), 8);
    v = (v << 32) | dec->read_bits(32);
    g.time[g.next] = v;
    g.last = g.next;
    g.diff[g.last] = 0;
    g.extreme[g.last] = 0;
  }
  void step() {
    if (g.diff[g.last] == 0) {
      const int multi = (int)dec->symbol(zero_model);
      if (multi == 1) {
        g.diff[g.last] = ic.decompress(0, 0);
        g.time[g.last] += (uint64_t)(int64_t)g.diff[g.last];
        g.extreme[g.last] = 0;
      } else if (multi == 2) {
        new_sequence();
      } else if (multi > 2) {
        g.last = (g.last + (uint32_t)multi - 2) & 3;
        step();
      }  // 0: unchanged
    } else {
      int multi = (int)dec->symbol(multi_model);
      const int32_t ld = g.diff[g.last];
      if (multi == 1) {
        g.time[g.last] += (uint64_t)(int64_t)ic.decompress(ld, 1);
        g.extreme[g.last] = 0;
      } else if (multi < kGps2Unchanged) {
        int32_t d;
        if (multi == 0) {
          d = ic.decompress(0, 7);
          g.bump(d);
        } else if (multi < kGpsMulti) {
          d = ic.decompress((int32_t)((uint32_t)multi * (uint32_t)ld), multi < 10 ? 2 : 3);
        } else if (multi == kGpsMulti) {
          d = ic.decompress((int32_t)((uint32_t)kGpsMulti * (uint32_t)ld), 4);
          g.bump(d);
        } else {
          multi = kGpsMulti - multi;
          if (multi > kGpsMultiMinus) {
            d = ic.decompress((int32_t)((uint32_t)multi * (uint32_t)ld), 5);
          } else {
            d = ic.decompress((int32_t)((uint32_t)kGpsMultiMinus * (uint32_t)ld), 6);
            g.bump(d);
          }
        }
        g.time[g.last] += (uint64_t)(int64_t)d;
      } else if (multi == kGps2CodeFull) {
        new_sequence();
      } else if (multi > kGps2CodeFull) {
        g.last = (g.last + (uint32_t)multi - (uint32_t)kGps2CodeFull) & 3;
        step();
      }  // kGps2Unchanged: unchanged
    }
  }
  void read(uint8_t* out) override {
    step();
    memcpy(out, &g.time[g.last], 8);
  }
};
struct GpsTime11V2Writer : ItemWriter {
  ArithmeticEncoder* enc;
  GpsState g;
  SymbolModel multi_model{(uint32_t)kGps2Total}, zero_model{6};
  IntegerCompressor ic;
  explicit GpsTime11V2Writer(ArithmeticEncoder* e) : enc(e), ic(e, 32, 9) {}
  void init(const uint8_t* first) override {
    g.init(first);
    multi_model.init();
    zero_model.init();
    ic.init();
  }
  void new_sequence(uint64_t t) {
    ic.compress((int32_t)(g.time[g.last] >> 32), (int32_t)(t >> 32), 8);
    enc->write_bits(32, (uint32_t)t);
    g.next = (g.next + 1) & 3;
    g.last = g.next;
    g.diff[g.last] = 0;
    g.extreme[g.last] = 0;
  }
  void write(const uint8_t* item) override {
    uint64_t t;
    memcpy(&t, item, 8);
    step(t);
  }
  void step(uint64_t t) {
    if (g.diff[g.last] == 0) {
      if (t == g.time[g.last]) {
        enc->symbol(zero_model, 0);  // unchanged
        return;
      }
      const int64_t d64 = (int64_t)(t - g.time[g.last]);
      const int32_t d32 = (int32_t)d64;
      if (d64 == (int64_t)d32) {
        enc->symbol(zero_model, 1);
        ic.compress(0, d32, 0);
        g.diff[g.last] = d32;
        g.extreme[g.last] = 0;
      } else {
        for (uint32_t i = 1; i < 4; ++i) {
          const int64_t o64 = (int64_t)(t - g.time[(g.last + i) & 3]);
          if (o64 == (int64_t)(int32_t)o64) {
            enc->symbol(zero_model, i + 2);
            g.last = (g.last + i) & 3;
            step(t);
            return;
          }
        }
        enc->symbol(zero_model, 2);
        new_sequence(t);
      }
      g.time[g.last] = t;
    } else {
      if (t == g.time[g.last]) {
        enc->symbol(multi_model, kGps2Unchanged);
        return;
      }
      const int64_t d64 = (int64_t)(t - g.time[g.last]);
      const int32_t d32 = (int32_t)d64;
      if (d64 == (int64_t)d32) {
        const int32_t ld = g.diff[g.last];
        const int32_t multi = Point14V3Writer::quantize((float)d32 / (float)ld);
        if (multi == 1) {
          enc->symbol(multi_model, 1);
          ic.compress(ld, d32, 1);
          g.extreme[g.last] = 0;
        } else if (multi > 0) {
          if (multi < kGpsMulti) {
            enc->symbol(multi_model, (uint32_t)multi);
            ic.compress((int32_t)((uint32_t)multi * (uint32_t)ld), d32, multi < 10 ? 2 : 3);
          } else {
            enc->symbol(multi_model, kGpsMulti);
            ic.compress((int32_t)((uint32_t)kGpsMulti * (uint32_t)ld), d32, 4);
            g.bump(d32);
          }
        } else if (multi < 0) {
          if (multi > kGpsMultiMinus) {
            enc->symbol(multi_model, (uint32_t)(kGpsMulti - multi));
            ic.compress((int32_t)((uint32_t)multi * (uint32_t)ld), d32, 5);
          } else {
            enc->symbol(multi_model, (uint32_t)(kGpsMulti - kGpsMultiMinus));
            ic.compress((int32_t)((uint32_t)kGpsMultiMinus * (uint32_t)ld), d32, 6);
            g.bump(d32);
          }
        } else {
          enc->symbol(multi_model, 0);
          ic.compress(0, d32, 7);
          g.bump(d32);
        }
      } else {
        for (uint32_t i = 1; i < 4; ++i) {
          const int64_t o64 = (int64_t)(t - g.time[(g.last + i) & 3]);
          if (o64 == (int64_t)(int32_t)o64) {
            enc->symbol(multi_model, (uint32_t)kGps2CodeFull + i);
            g.last = (g.last + i) & 3;
            step(t);
            return;
          }
        }
        enc->symbol(multi_model, kGps2CodeFull);
        new_sequence(t);
      }
      g.time[g.last] = t;
    }
  }
};


int check_items(const char* who, const uint16_t* items, int nitems, int point_size) {
  int total = 0;
  for (int i = 0; i < nitems; ++i) {
    const int type = items[3 * i], size = items[3 * i + 1], version = items[3 * i + 2];
    const bool ok = (type == 6 && size == 20 && version == 2) || (type == 7 && size == 8 && version == 2) || (type == 8 && size == 6 && version == 2) ||
                    (type == 0 && size > 0 && version == 2);
    if (!ok)
      return vapc_set_error(VAPC_E_INVALID, "%s: LASzip item type %d (size %d, version %d) is not supported (POINT10 v2, GPSTIME11 v2, RGB12 v2 and BYTE v2 are)", who, type,
                            size, version);
    total += size;
  }
  if (total != point_size) return vapc_set_error(VAPC_E_INVALID, "%s: item sizes (%d) do not add up to the point size (%d)", who, total, point_size);
  return VAPC_OK;
}
}  // namespace

// Decode `npoints` records of `point_size` bytes into records_out_host.  items_host: nitems triples (type, size,
// version) from the laszip VLR (user id "laszip encoded", record id 22204).  Everything is host memory.
extern "C" int vapc_laz_decode(const uint8_t* file_host, int64_t file_bytes, int64_t offset_to_points, int64_t npoints, int32_t point_size,
                               int32_t compressor, uint32_t chunk_size, const uint16_t* items_host, int32_t nitems, uint8_t* records_out_host) {
  if (!file_host || !items_host || !records_out_host || file_bytes < 0 || npoints < 0 || nitems < 1)
    return vapc_set_error(VAPC_E_INVALID, "vapc_laz_decode: bad arguments");
  if (compressor == 3) return decode_layered(file_host, file_bytes, offset_to_points, npoints, point_size, chunk_size, items_host, nitems, records_out_host);
  if (compressor != 2)
    return vapc_set_error(VAPC_E_INVALID, "vapc_laz_decode: LASzip compressor %d is not supported (2 = pointwise chunked and 3 = layered chunked are)", compressor);
  if (const int rc = check_items("vapc_laz_decode", items_host, nitems, point_size)) return rc;
  if (npoints == 0) return VAPC_OK;
  if (chunk_size == 0 || chunk_size == 0xFFFFFFFFu) return vapc_set_error(VAPC_E_INVALID, "vapc_laz_decode: variable chunk sizes are not supported");
  if (offset_to_points + 8 > file_bytes) return vapc_set_error(VAPC_E_INVALID, "vapc_laz_decode: truncated file");
  int64_t table_pos;
  memcpy(&table_pos, file_host + offset_to_points, 8);
  if (table_pos == -1) memcpy(&table_pos, file_host + file_bytes - 8, 8);  // written at the end when the stream was not seekable
  const int64_t nchunks = (npoints + chunk_size - 1) / chunk_size;
  std::vector<int64_t> chunk_start((size_t)nchunks + 1);
  chunk_start[0] = offset_to_points + 8;
  if (nchunks > 1) {
    // chunk table: u32 version, u32 count, then the arithmetic-coded byte sizes of the chunks
    if (table_pos < chunk_start[0] || table_pos + 8 > file_bytes) return vapc_set_error(VAPC_E_INVALID, "vapc_laz_decode: chunk table missing");
    uint32_t version, count;
    memcpy(&version, file_host + table_pos, 4);
    memcpy(&count, file_host + table_pos + 4, 4);
    if (version != 0 || (int64_t)count < nchunks) return vapc_set_error(VAPC_E_INVALID, "vapc_laz_decode: unexpected chunk table (version %u, %u chunks)", version, count);
    ByteReader br{file_host + table_pos + 8, file_host + file_bytes};
    ArithmeticDecoder dec;
    dec.init(&br);
    IntegerDecompressor ic(&dec, 32, 2);
    int32_t prev = 0;
    for (int64_t i = 0; i < nchunks; ++i) {
      prev = ic.decompress(prev, 1);
      chunk_start[(size_t)i + 1] = chunk_start[(size_t)i] + (uint32_t)prev;
    }
    if (br.overrun) return vapc_set_error(VAPC_E_INVALID, "vapc_laz_decode: chunk table truncated");
  }
  struct State {
    ArithmeticDecoder dec;
    std::vector<std::unique_ptr<ItemReader>> readers;
  };
  std::vector<int> sizes;
  for (int i = 0; i < nitems; ++i) sizes.push_back(items_host[3 * i + 1]);
  auto make = [&]() {
    std::unique_ptr<State> st(new State());
    for (int i = 0; i < nitems; ++i) {
      if (items_host[3 * i] == 6) st->readers.emplace_back(new Point10V2(&st->dec));
      else if (items_host[3 * i] == 7) st->readers.emplace_back(new GpsTime11V2(&st->dec));
      else if (items_host[3 * i] == 8) st->readers.emplace_back(new Rgb12V2(&st->dec));
      else st->readers.emplace_back(new ByteV2(&st->dec, items_host[3 * i + 1]));
    }
    return st;
  };
  const bool has_table = table_pos >= chunk_start[0] && table_pos + 8 <= file_bytes;
  auto one_chunk = [&](int64_t c, State& st) -> int {
    const int64_t first = c * (int64_t)chunk_size;
    const int64_t count = std::min<int64_t>(chunk_size, npoints - first);
    const int64_t begin = chunk_start[(size_t)c];
    const int64_t limit = nchunks > 1 && c + 1 < nchunks ? chunk_start[(size_t)c + 1] : (has_table ? table_pos : file_bytes);
    if (begin + point_size > file_bytes || limit > file_bytes) return vapc_set_error(VAPC_E_INVALID, "vapc_laz_decode: chunk %lld outside the file", (long long)c);
    uint8_t* out = records_out_host + first * point_size;
    memcpy(out, file_host + begin, (size_t)point_size);  // the first point of a chunk is stored raw
    int off = 0;
    for (size_t i = 0; i < st.readers.size(); ++i) {
      st.readers[i]->init(out + off);
      off += sizes[i];
    }
    if (count > 1) {
      ByteReader br{file_host + begin + point_size, file_host + limit};
      st.dec.init(&br);
      for (int64_t p = 1; p < count; ++p) {
        uint8_t* rec = out + p * point_size;
        off = 0;
        for (size_t i = 0; i < st.readers.size(); ++i) {
          st.readers[i]->read(rec + off);
          off += sizes[i];
        }
      }
      // a correct decode ends on the last byte of the chunk (the chunk table marks the end of the last one)
      if (br.overrun || ((nchunks > 1 || has_table) && br.p != br.end))
        return vapc_set_error(VAPC_E_INVALID, "vapc_laz_decode: chunk %lld is truncated or corrupt", (long long)c);
    }
    return VAPC_OK;
  };
  return parallel_chunks(nchunks, make, one_chunk);
}

// Encode `npoints` records into a LASzip "pointwise chunked" stream (compressor 2): [int64 position of the chunk table,
// relative to the start of the stream + stream_offset_host][chunks: first point raw, the rest arithmetic-coded]
// [chunk table: u32 version 0, u32 count, arithmetic-coded chunk byte sizes].  stream_offset_host = file position at which the
// caller will place the stream (= the header's offset to point data).  Returns the stream length in *out_bytes_host;
// VAPC_E_WORKSPACE (with the needed size in *out_bytes_host) when out_cap is too small.
extern "C" int vapc_laz_encode(const uint8_t* records_host, int64_t npoints, int32_t point_size, uint32_t chunk_size, const uint16_t* items_host,
                               int32_t nitems, int64_t stream_offset_host, uint8_t* out_host, int64_t out_cap, int64_t* out_bytes_host) {
  if (!items_host || !out_bytes_host || npoints < 0 || nitems < 1 || point_size < 1 || (npoints > 0 && !records_host))
    return vapc_set_error(VAPC_E_INVALID, "vapc_laz_encode: bad arguments");
  if (chunk_size == 0 || chunk_size == 0xFFFFFFFFu) return vapc_set_error(VAPC_E_INVALID, "vapc_laz_encode: a fixed chunk size is required");
  if (items_host[0] == 10) {  // LAS 1.4: layered chunked (compressor 3), POINT14 v3 (+ BYTE14 v3)
    int extra = 0, rgb = 0;
    if (const int rc = layered_items("vapc_laz_encode", items_host, nitems, point_size, &rgb, &extra)) return rc;
    std::vector<uint8_t> lay(8, 0);
    std::vector<uint32_t> sizes;
    if (const int rc = encode_layered(records_host, npoints, point_size, chunk_size, rgb, extra, lay, sizes)) return rc;
    const int64_t tpos = stream_offset_host + (int64_t)lay.size();
    memcpy(lay.data(), &tpos, 8);
    write_chunk_table(lay, sizes);
    *out_bytes_host = (int64_t)lay.size();
    if (!out_host || out_cap < (int64_t)lay.size()) return vapc_set_error(VAPC_E_WORKSPACE, "vapc_laz_encode: output buffer too small (%lld bytes needed)", (long long)lay.size());
    memcpy(out_host, lay.data(), lay.size());
    return VAPC_OK;
  }
  if (const int rc = check_items("vapc_laz_encode", items_host, nitems, point_size)) return rc;
  std::vector<uint8_t> out(8, 0);
  struct State {
    ArithmeticEncoder enc;
    std::vector<std::unique_ptr<ItemWriter>> writers;
  };
  std::vector<int> sizes;
  for (int i = 0; i < nitems; ++i) sizes.push_back(items_host[3 * i + 1]);
  auto make = [&]() {
    std::unique_ptr<State> st(new State());
    for (int i = 0; i < nitems; ++i) {
      if (items_host[3 * i] == 6) st->writers.emplace_back(new Point10V2Writer(&st->enc));
      else if (items_host[3 * i] == 7) st->writers.emplace_back(new GpsTime11V2Writer(&st->enc));
      else if (items_host[3 * i] == 8) st->writers.emplace_back(new Rgb12V2Writer(&st->enc));
      else st->writers.emplace_back(new ByteV2Writer(&st->enc, items_host[3 * i + 1]));
    }
    return st;
  };
  const int64_t nchunks = (npoints + chunk_size - 1) / chunk_size;
  std::vector<std::vector<uint8_t>> parts((size_t)nchunks);
  auto one_chunk = [&](int64_t c, State& st) -> int {
    std::vector<uint8_t>& o = parts[(size_t)c];
    const int64_t first = c * (int64_t)chunk_size;
    const int64_t count = std::min<int64_t>(chunk_size, npoints - first);
    const uint8_t* rec = records_host + first * point_size;
    o.insert(o.end(), rec, rec + point_size);  // the first point of a chunk is stored raw
    int off = 0;
    for (size_t i = 0; i < st.writers.size(); ++i) {
      st.writers[i]->init(rec + off);
      off += sizes[i];
    }
    st.enc.init(&o);
    for (int64_t p = 1; p < count; ++p) {
      const uint8_t* r = rec + p * point_size;
      off = 0;
      for (size_t i = 0; i < st.writers.size(); ++i) {
        st.writers[i]->write(r + off);
        off += sizes[i];
      }
    }
    st.enc.done();
    return VAPC_OK;
  };
  if (const int rc = parallel_chunks(nchunks, make, one_chunk)) return rc;
  std::vector<uint32_t> chunk_bytes;
  for (auto& part : parts) {
    out.insert(out.end(), part.begin(), part.end());
    chunk_bytes.push_back((uint32_t)part.size());
    std::vector<uint8_t>().swap(part);
  }
  const int64_t table_pos = stream_offset_host + (int64_t)out.size();
  memcpy(out.data(), &table_pos, 8);
  write_chunk_table(out, chunk_bytes);
  *out_bytes_host = (int64_t)out.size();
  if (!out_host || out_cap < (int64_t)out.size()) return vapc_set_error(VAPC_E_WORKSPACE, "vapc_laz_encode: output buffer too small (%lld bytes needed)", (long long)out.size());
  memcpy(out_host, out.data(), out.size());
  return VAPC_OK;
}
