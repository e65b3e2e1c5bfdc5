// NIfTI-1 single-file (.nii / .nii.gz) reader and writer: what the tools read and write
// through itk::ImageFileReader / ImageFileWriter in the reference (OUT_FILE_TYPE ".nii.gz",
// tools/ExtractFeatures.cxx:15).  Voxel data are cast to the requested pixel type on read
// (scl_slope / scl_inter applied as ITK's NIfTI IO does); spacing comes from pixdim.  The
// source header travels with the image so that outputs keep the input's orientation.
#ifndef IFE_B200_NIFTI_IO_H
#define IFE_B200_NIFTI_IO_H
#include <zlib.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "ife/Image.h"

namespace ife {
namespace nifti {

enum { DT_UINT8 = 2, DT_INT16 = 4, DT_INT32 = 8, DT_FLOAT32 = 16, DT_FLOAT64 = 64, DT_INT8 = 256,
       DT_UINT16 = 512, DT_UINT32 = 768 };

inline void swap_bytes(void* p, size_t size, size_t count) {
  unsigned char* b = static_cast<unsigned char*>(p);
  for (size_t i = 0; i < count; ++i, b += size)
    for (size_t j = 0; j < size / 2; ++j) std::swap(b[j], b[size - 1 - j]);
}

template <typename T> T rd(const std::vector<unsigned char>& h, size_t off, bool swapped) {
  T v;
  std::memcpy(&v, h.data() + off, sizeof(T));
  if (swapped) swap_bytes(&v, sizeof(T), 1);
  return v;
}
template <typename T> void wr(std::vector<unsigned char>& h, size_t off, T v) { std::memcpy(h.data() + off, &v, sizeof(T)); }

inline bool ends_with(const std::string& s, const std::string& e) {
  return s.size() >= e.size() && s.compare(s.size() - e.size(), e.size(), e) == 0;
}

template <typename TOut, typename TIn>
void convert(const std::vector<unsigned char>& raw, bool swapped, double slope, double inter, TOut* out, size_t n) {
  const bool scale = slope != 0.0 && !(slope == 1.0 && inter == 0.0);
  for (size_t i = 0; i < n; ++i) {
    TIn v;
    std::memcpy(&v, raw.data() + i * sizeof(TIn), sizeof(TIn));
    if (swapped) swap_bytes(&v, sizeof(TIn), 1);
    out[i] = scale ? static_cast<TOut>((double)v * slope + inter) : static_cast<TOut>(v);
  }
}

template <typename T>
typename Image<T>::Pointer Read(const std::string& path) {
  gzFile f = gzopen(path.c_str(), "rb");
  if (!f) throw std::runtime_error("cannot open '" + path + "'");
  std::vector<unsigned char> hdr(352, 0);
  if (gzread(f, hdr.data(), 348) != 348) { gzclose(f); throw std::runtime_error("'" + path + "': short NIfTI header"); }
  int sizeof_hdr;
  std::memcpy(&sizeof_hdr, hdr.data(), 4);
  bool swapped = false;
  if (sizeof_hdr != 348) {
    swap_bytes(&sizeof_hdr, 4, 1);
    if (sizeof_hdr != 348) { gzclose(f); throw std::runtime_error("'" + path + "' is not a NIfTI-1 file"); }
    swapped = true;
  }
  if (std::memcmp(hdr.data() + 344, "n+1", 3) != 0) { gzclose(f); throw std::runtime_error("'" + path + "': only single-file NIfTI-1 (n+1) is supported"); }
  const int ndim = rd<int16_t>(hdr, 40, swapped);
  int dims[3] = {1, 1, 1};
  for (int d = 0; d < 3 && d < ndim; ++d) dims[d] = rd<int16_t>(hdr, 42 + 2 * d, swapped);
  for (int d = 3; d < ndim && d < 7; ++d)
    if (rd<int16_t>(hdr, 42 + 2 * d, swapped) > 1) { gzclose(f); throw std::runtime_error("'" + path + "': more than 3 dimensions"); }
  const int datatype = rd<int16_t>(hdr, 70, swapped);
  const float vox_offset = rd<float>(hdr, 108, swapped);
  const double slope = rd<float>(hdr, 112, swapped), inter = rd<float>(hdr, 116, swapped);
  auto img = Image<T>::New();
  Geometry g;
  g.size = {{dims[0], dims[1], dims[2]}};
  for (int d = 0; d < 3; ++d) {
    const double p = std::fabs((double)rd<float>(hdr, 76 + 4 * (d + 1), swapped));
    g.spacing[d] = p > 0 ? p : 1.0;
  }
  g.origin = {{(double)rd<float>(hdr, 268, swapped), (double)rd<float>(hdr, 272, swapped), (double)rd<float>(hdr, 276, swapped)}};
  if (!swapped) g.nifti_header.assign(hdr.begin(), hdr.begin() + 348);
  img->SetGeometry(g);
  img->Allocate();
  size_t bpv;
  switch (datatype) {
    case DT_UINT8: case DT_INT8: bpv = 1; break;
    case DT_INT16: case DT_UINT16: bpv = 2; break;
    case DT_INT32: case DT_UINT32: case DT_FLOAT32: bpv = 4; break;
    case DT_FLOAT64: bpv = 8; break;
    default: gzclose(f); throw std::runtime_error("'" + path + "': unsupported NIfTI datatype " + std::to_string(datatype));
  }
  const long skip = (long)(vox_offset > 348 ? vox_offset : 352) - 348;
  std::vector<unsigned char> junk(skip > 0 ? skip : 1);
  if (skip > 0 && gzread(f, junk.data(), (unsigned)skip) != skip) { gzclose(f); throw std::runtime_error("'" + path + "': truncated"); }
  const size_t n = g.voxels();
  std::vector<unsigned char> raw(n * bpv);
  size_t got = 0;
  while (got < raw.size()) {
    const unsigned want = (unsigned)std::min<size_t>(raw.size() - got, 1u << 30);
    const int r = gzread(f, raw.data() + got, want);
    if (r <= 0) break;
    got += (size_t)r;
  }
  gzclose(f);
  if (got != raw.size()) throw std::runtime_error("'" + path + "': truncated voxel data");
  T* out = img->GetBufferPointer();
  switch (datatype) {
    case DT_UINT8: convert<T, uint8_t>(raw, swapped, slope, inter, out, n); break;
    case DT_INT8: convert<T, int8_t>(raw, swapped, slope, inter, out, n); break;
    case DT_INT16: convert<T, int16_t>(raw, swapped, slope, inter, out, n); break;
    case DT_UINT16: convert<T, uint16_t>(raw, swapped, slope, inter, out, n); break;
    case DT_INT32: convert<T, int32_t>(raw, swapped, slope, inter, out, n); break;
    case DT_UINT32: convert<T, uint32_t>(raw, swapped, slope, inter, out, n); break;
    case DT_FLOAT32: convert<T, float>(raw, swapped, slope, inter, out, n); break;
    case DT_FLOAT64: convert<T, double>(raw, swapped, slope, inter, out, n); break;
  }
  return img;
}

template <typename T> struct DataType;
template <> struct DataType<float> { static const int16_t code = DT_FLOAT32; };
template <> struct DataType<unsigned char> { static const int16_t code = DT_UINT8; };
template <> struct DataType<short> { static const int16_t code = DT_INT16; };
template <> struct DataType<unsigned short> { static const int16_t code = DT_UINT16; };

template <typename T>
void Write(const std::string& path, const Geometry& g, const T* data) {
  std::vector<unsigned char> hdr(352, 0);
  if (g.nifti_header.size() == 348) std::copy(g.nifti_header.begin(), g.nifti_header.end(), hdr.begin());
  wr<int32_t>(hdr, 0, 348);
  int16_t dim[8] = {3, (int16_t)g.size[0], (int16_t)g.size[1], (int16_t)g.size[2], 1, 1, 1, 1};
  std::memcpy(hdr.data() + 40, dim, sizeof(dim));
  wr<int16_t>(hdr, 70, DataType<T>::code);
  wr<int16_t>(hdr, 72, (int16_t)(8 * sizeof(T)));
  if (g.nifti_header.size() != 348) {
    float pixdim[8] = {1, (float)g.spacing[0], (float)g.spacing[1], (float)g.spacing[2], 0, 0, 0, 0};
    std::memcpy(hdr.data() + 76, pixdim, sizeof(pixdim));
    wr<int16_t>(hdr, 254, 1);  // sform_code: scanner coordinates, axis-aligned
    float sx[4] = {(float)g.spacing[0], 0, 0, (float)g.origin[0]};
    float sy[4] = {0, (float)g.spacing[1], 0, (float)g.origin[1]};
    float sz[4] = {0, 0, (float)g.spacing[2], (float)g.origin[2]};
    std::memcpy(hdr.data() + 280, sx, 16);
    std::memcpy(hdr.data() + 296, sy, 16);
    std::memcpy(hdr.data() + 312, sz, 16);
    hdr[123] = 2;  // xyzt_units: millimetres
  }
  wr<float>(hdr, 108, 352.0f);
  wr<float>(hdr, 112, 1.0f);   // scl_slope
  wr<float>(hdr, 116, 0.0f);   // scl_inter
  std::memcpy(hdr.data() + 344, "n+1", 4);
  const size_t bytes = g.voxels() * sizeof(T);
  if (ends_with(path, ".gz")) {
    gzFile f = gzopen(path.c_str(), "wb1");
    if (!f) throw std::runtime_error("cannot write '" + path + "'");
    bool ok = gzwrite(f, hdr.data(), 352) == 352;
    size_t done = 0;
    while (ok && done < bytes) {
      const unsigned want = (unsigned)std::min<size_t>(bytes - done, 1u << 30);
      const int w = gzwrite(f, reinterpret_cast<const unsigned char*>(data) + done, want);
      ok = w > 0;
      done += ok ? (size_t)w : 0;
    }
    ok = (gzclose(f) == Z_OK) && ok;
    if (!ok) throw std::runtime_error("error writing '" + path + "'");
  } else {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot write '" + path + "'");
    bool ok = std::fwrite(hdr.data(), 1, 352, f) == 352 && std::fwrite(data, 1, bytes, f) == bytes;
    ok = (std::fclose(f) == 0) && ok;
    if (!ok) throw std::runtime_error("error writing '" + path + "'");
  }
}

template <typename T>
void Write(const std::string& path, const Image<T>& img) { Write(path, img.GetGeometry(), img.GetBufferPointer()); }

}  // namespace nifti
}  // namespace ife
#endif
