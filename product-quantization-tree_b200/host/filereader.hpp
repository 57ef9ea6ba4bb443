// filereader.hpp -- reader/writer of the reference's .umem/.imem containers
// (reader: utils/filereader.hpp:8-131; writer: convert/filehelper.hpp:251-278).
// Layout: ASCII "<N>\n<D>\n", zero padding up to byte 20, payload from byte 20.
// FileReader<float> reads a uint8 payload and widens it to float like the reference.
#ifndef PQT_B200_HOST_FILEREADER_HPP
#define PQT_B200_HOST_FILEREADER_HPP

#include <cstdint>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace pqt_io {

constexpr std::streamoff kPayloadOffset = 20;

inline void readHeader(const std::string& path, uint32_t& n, uint32_t& d) {
  std::ifstream f(path, std::ios::in | std::ios::binary);
  if (!f.good()) throw std::runtime_error("cannot open file " + path);
  f >> n >> d;
  if (f.fail()) throw std::runtime_error("bad header in " + path);
}

template <typename Stored>
std::vector<Stored> readPayload(const std::string& path, uint32_t d, size_t num, size_t offset) {
  std::ifstream f(path, std::ios::in | std::ios::binary);
  if (!f.good()) throw std::runtime_error("cannot open file " + path);
  std::vector<Stored> buf(num * d);
  f.seekg(kPayloadOffset + (std::streamoff)(sizeof(Stored) * offset * d), std::ios::beg);
  f.read(reinterpret_cast<char*>(buf.data()), (std::streamsize)(buf.size() * sizeof(Stored)));
  if ((size_t)f.gcount() != buf.size() * sizeof(Stored))
    throw std::runtime_error("short read in " + path);
  return buf;
}

template <typename T>
void writeMem(const std::string& path, const T* data, uint32_t n, uint32_t d) {
  std::ofstream f(path, std::ios::out | std::ios::binary);
  if (!f.good()) throw std::runtime_error("cannot open " + path + " for writing");
  std::string hdr = std::to_string(n) + "\n" + std::to_string(d) + "\n";
  if (hdr.size() > (size_t)kPayloadOffset) throw std::runtime_error("header too long");
  hdr.resize((size_t)kPayloadOffset, '\0');
  f.write(hdr.data(), (std::streamsize)hdr.size());
  f.write(reinterpret_cast<const char*>(data), (std::streamsize)((size_t)n * d * sizeof(T)));
}

}  // namespace pqt_io

template <typename T>
class FileReader {  // uint8 payload -> T
 public:
  explicit FileReader(const std::string& fs) : filename_(fs) { pqt_io::readHeader(fs, n_, d_); }
  uint32_t num() const { return n_; }
  uint32_t dim() const { return d_; }
  std::vector<T> data() { return data(n_, 0); }
  std::vector<T> data(size_t num, size_t offset = 0) {
    std::vector<uint8_t> raw = pqt_io::readPayload<uint8_t>(filename_, d_, num, offset);
    return std::vector<T>(raw.begin(), raw.end());
  }

 private:
  std::string filename_;
  uint32_t n_ = 0, d_ = 0;
};

template <>
class FileReader<int> {  // int32 payload (.imem)
 public:
  explicit FileReader(const std::string& fs) : filename_(fs) { pqt_io::readHeader(fs, n_, d_); }
  uint32_t entries() const { return n_; }
  uint32_t dimension() const { return d_; }
  std::vector<int> data() { return data(n_, 0); }
  std::vector<int> data(size_t num, size_t offset) {
    return pqt_io::readPayload<int>(filename_, d_, num, offset);
  }

 private:
  std::string filename_;
  uint32_t n_ = 0, d_ = 0;
};

#endif
