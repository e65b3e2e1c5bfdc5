// Minimal 3-D image containers for the host side of the B200 path.  The reference works on
// itk::Image<float,3> / itk::Image<unsigned char,3> / itk::VectorImage<float,3>
// (tools/ExtractFeatures.cxx:81-86); ITK is not a dependency of this library, so these
// carry exactly what the hot path needs: size, spacing, origin, an x-fastest buffer.
//
// Storage is page-locked (ife_cuda_host_alloc) whenever a CUDA device is present, so that
// the library's host<->device copies of an image run asynchronously at full PCIe rate, and it
// is never value-initialised (a 512x512x400 scan with four scales is 13 GB of results).
// Images can be non-owning VIEWS of another image's storage (one component of a
// VectorImage, one scale of a multi-scale result), and carry the "source" hook that gives
// ITK's pull semantics: consumer->Update() first updates whatever produces its inputs.
#ifndef IFE_B200_IMAGE_H
#define IFE_B200_IMAGE_H
#include <algorithm>
#include <array>
#include <chrono>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <new>
#include <stdexcept>
#include <vector>

#include "ife_cuda.h"

namespace ife {

struct Geometry {
  std::array<int, 3> size{{0, 0, 0}};          // nx, ny, nz
  std::array<double, 3> spacing{{1.0, 1.0, 1.0}};
  std::array<double, 3> origin{{0.0, 0.0, 0.0}};
  std::vector<unsigned char> nifti_header;     // raw header of the file it came from (for writing)
  size_t voxels() const { return (size_t)size[0] * size[1] * size[2]; }
};

// seconds spent allocating (page-locking) pixel storage so far in this process
inline double& alloc_seconds() { static double s = 0.0; return s; }
inline std::mutex& alloc_seconds_mutex() { static std::mutex mu; return mu; }   // images are read (and allocated) concurrently

// Contiguous host storage: page-locked when the library can provide it, pageable otherwise
// (no GPU: the IO classes still work).  Growth leaves new elements uninitialised.
template <typename T>
class PixelContainer {
public:
  typedef T value_type;
  typedef T* iterator;
  typedef const T* const_iterator;
  PixelContainer() {}
  ~PixelContainer() { release(); }
  PixelContainer(const PixelContainer&) = delete;
  PixelContainer& operator=(const PixelContainer&) = delete;

  void resize(size_t n) {
    if (n == m_Size) return;
    if (n > m_Capacity) {
      release();
      if (n) {
        void* p = nullptr;
        const auto t0 = std::chrono::steady_clock::now();
        const int rc = ife_cuda_host_alloc(n * sizeof(T), &p);
        {
          std::lock_guard<std::mutex> lk(alloc_seconds_mutex());
          alloc_seconds() += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        }
        if (rc == IFE_OK && p) {
          m_Pinned = true;
        } else {
          p = std::malloc(n * sizeof(T));
          if (!p) throw std::bad_alloc();
          m_Pinned = false;
        }
        m_Data = static_cast<T*>(p);
        m_Capacity = n;
      }
    }
    m_Size = n;
  }
  void assign(size_t n, const T& v) { resize(n); std::fill(m_Data, m_Data + n, v); }
  template <typename It>
  void assign(It first, It last) {
    resize((size_t)std::distance(first, last));
    std::copy(first, last, m_Data);
  }
  T* data() { return m_Data; }
  const T* data() const { return m_Data; }
  size_t size() const { return m_Size; }
  bool pinned() const { return m_Pinned; }
  iterator begin() { return m_Data; }
  iterator end() { return m_Data + m_Size; }
  const_iterator begin() const { return m_Data; }
  const_iterator end() const { return m_Data + m_Size; }
  T& operator[](size_t i) { return m_Data[i]; }
  const T& operator[](size_t i) const { return m_Data[i]; }
  bool operator==(const PixelContainer& o) const {
    return m_Size == o.m_Size && (m_Size == 0 || std::equal(begin(), end(), o.begin()));
  }

private:
  void release() {
    if (m_Data) {
      if (m_Pinned) ife_cuda_host_free(m_Data);
      else std::free(m_Data);
    }
    m_Data = nullptr;
    m_Size = m_Capacity = 0;
  }
  T* m_Data = nullptr;
  size_t m_Size = 0, m_Capacity = 0;
  bool m_Pinned = false;
};

// What an image and a view of it share.
template <typename T>
struct PixelStore {
  PixelContainer<T> data;
};

template <typename T>
class Image {
public:
  typedef T PixelType;
  typedef std::shared_ptr<Image> Pointer;
  typedef PixelContainer<T> PixelContainerType;
  static Pointer New() { return std::make_shared<Image>(); }

  void SetRegions(int nx, int ny, int nz) { m_Geom.size = {{nx, ny, nz}}; }
  void SetSpacing(double sx, double sy, double sz) { m_Geom.spacing = {{sx, sy, sz}}; }
  // ITK's Allocate() leaves the buffer uninitialised too; Allocate(true) zero-fills
  void Allocate(bool initialize = false) {
    own().data.resize(m_Geom.voxels());
    m_Ptr = m_Store->data.data();
    m_N = m_Geom.voxels();
    if (initialize && m_N) std::memset(m_Ptr, 0, m_N * sizeof(T));
  }
  const Geometry& GetGeometry() const { return m_Geom; }
  void SetGeometry(const Geometry& g) { m_Geom = g; }
  const std::array<int, 3>& GetSize() const { return m_Geom.size; }
  const std::array<double, 3>& GetSpacing() const { return m_Geom.spacing; }
  T* GetBufferPointer() { return m_Ptr; }
  const T* GetBufferPointer() const { return m_Ptr; }
  // the owned storage (an image that is a view has none of its own: it gets one, detached)
  PixelContainerType& GetPixelContainer() {
    PixelContainerType& c = own().data;
    return c;
  }
  // after the container was resized / assigned directly (the IO classes do)
  void SyncWithContainer() { m_Ptr = m_Store->data.data(); m_N = m_Store->data.size(); }
  size_t GetNumberOfPixels() const { return m_N; }
  T& GetPixel(int x, int y, int z) { return m_Ptr[index(x, y, z)]; }
  const T& GetPixel(int x, int y, int z) const { return m_Ptr[index(x, y, z)]; }
  bool SameBufferContent(const Image& o) const {
    return m_N == o.m_N && (m_N == 0 || std::equal(m_Ptr, m_Ptr + m_N, o.m_Ptr));
  }

  // Non-owning view of `n` elements at `p` inside `keep`'s storage.
  void SetView(const Geometry& g, T* p, size_t n, std::shared_ptr<void> keep) {
    m_Geom = g; m_Ptr = p; m_N = n; m_Keep = std::move(keep); m_Store.reset();
  }

  // ITK pull semantics: whoever produces this image (a reader, a filter) registers itself here
  // and consumers call UpdateSource() before they read the buffer (itk::DataObject::Update).
  void SetSource(std::function<void()> f) { m_Source = std::move(f); }
  void UpdateSource() const { if (m_Source) m_Source(); }
  void Update() const { UpdateSource(); }

private:
  size_t index(int x, int y, int z) const { return (size_t)x + (size_t)m_Geom.size[0] * (y + (size_t)m_Geom.size[1] * z); }
  PixelStore<T>& own() {
    if (!m_Store) { m_Store = std::make_shared<PixelStore<T> >(); m_Keep.reset(); m_Ptr = nullptr; m_N = 0; }
    return *m_Store;
  }
  Geometry m_Geom;
  std::shared_ptr<PixelStore<T> > m_Store;
  std::shared_ptr<void> m_Keep;
  T* m_Ptr = nullptr;
  size_t m_N = 0;
  std::function<void()> m_Source;
};

// Multi-component image stored as SoA planes (component k = one contiguous volume), which
// is what the GPU writes and what every tool consumes (one file per component).
template <typename T>
class VectorImage {
public:
  typedef T InternalPixelType;
  typedef std::shared_ptr<VectorImage> Pointer;
  static Pointer New() { return std::make_shared<VectorImage>(); }
  void SetGeometry(const Geometry& g) { m_Geom = g; }
  const Geometry& GetGeometry() const { return m_Geom; }
  void SetNumberOfComponentsPerPixel(unsigned n) { m_Components = n; }
  unsigned GetNumberOfComponentsPerPixel() const { return m_Components; }
  void Allocate(bool initialize = false) {
    if (!m_Store) m_Store = std::make_shared<PixelStore<T> >();
    m_Store->data.resize(m_Geom.voxels() * m_Components);
    m_Ptr = m_Store->data.data();
    m_Keep.reset();
    if (initialize && m_Ptr) std::memset(m_Ptr, 0, m_Geom.voxels() * m_Components * sizeof(T));
  }
  T* GetBufferPointer() { return m_Ptr; }
  const T* GetBufferPointer() const { return m_Ptr; }
  const T* GetComponentPointer(unsigned k) const { return m_Ptr + (size_t)k * m_Geom.voxels(); }
  // component k as an image that shares this image's storage (no copy):
  // itk::VectorIndexSelectionCastImageFilter without the cast
  typename Image<T>::Pointer GetComponentView(unsigned k) const {
    if (k >= m_Components) throw std::out_of_range("component index");
    auto img = Image<T>::New();
    img->SetView(m_Geom, m_Ptr + (size_t)k * m_Geom.voxels(), m_Geom.voxels(), keepalive());
    return img;
  }
  // a copy of component k
  typename Image<T>::Pointer ExtractComponent(unsigned k) const {
    if (k >= m_Components) throw std::out_of_range("component index");
    auto img = Image<T>::New();
    img->SetGeometry(m_Geom);
    img->GetPixelContainer().assign(GetComponentPointer(k), GetComponentPointer(k) + m_Geom.voxels());
    img->SyncWithContainer();
    return img;
  }
  std::vector<T> GetPixel(int x, int y, int z) const {
    std::vector<T> p(m_Components);
    const size_t i = (size_t)x + (size_t)m_Geom.size[0] * (y + (size_t)m_Geom.size[1] * z);
    for (unsigned k = 0; k < m_Components; ++k) p[k] = m_Ptr[(size_t)k * m_Geom.voxels() + i];
    return p;
  }
  // Non-owning view (one scale of a multi-scale result).
  void SetView(const Geometry& g, unsigned components, T* p, std::shared_ptr<void> keep) {
    m_Geom = g; m_Components = components; m_Ptr = p; m_Keep = std::move(keep); m_Store.reset();
  }
  std::shared_ptr<void> keepalive() const { return m_Store ? std::shared_ptr<void>(m_Store) : m_Keep; }

  void SetSource(std::function<void()> f) { m_Source = std::move(f); }
  void UpdateSource() const { if (m_Source) m_Source(); }
  void Update() const { UpdateSource(); }

private:
  Geometry m_Geom;
  unsigned m_Components = 0;
  std::shared_ptr<PixelStore<T> > m_Store;
  std::shared_ptr<void> m_Keep;
  T* m_Ptr = nullptr;
  std::function<void()> m_Source;
};

}  // namespace ife
#endif
