// Minimal 3-D image containers for the host side of the B200 path.  The reference works on
// itk::Image<float,3> / itk::Image<unsigned char,3> / itk::VectorImage<float,3>
// (tools/ExtractFeatures.cxx:81-86); ITK is not a dependency of this library, so these
// carry exactly what the hot path needs: size, spacing, origin, an x-fastest buffer.
#ifndef IFE_B200_IMAGE_H
#define IFE_B200_IMAGE_H
#include <array>
#include <cstddef>
#include <memory>
#include <stdexcept>
#include <vector>

namespace ife {

struct Geometry {
  std::array<int, 3> size{{0, 0, 0}};          // nx, ny, nz
  std::array<double, 3> spacing{{1.0, 1.0, 1.0}};
  std::array<double, 3> origin{{0.0, 0.0, 0.0}};
  std::vector<unsigned char> nifti_header;     // raw header of the file it came from (for writing)
  size_t voxels() const { return (size_t)size[0] * size[1] * size[2]; }
};

template <typename T>
class Image {
public:
  typedef T PixelType;
  typedef std::shared_ptr<Image> Pointer;
  static Pointer New() { return std::make_shared<Image>(); }

  void SetRegions(int nx, int ny, int nz) { m_Geom.size = {{nx, ny, nz}}; }
  void SetSpacing(double sx, double sy, double sz) { m_Geom.spacing = {{sx, sy, sz}}; }
  void Allocate() { m_Data.assign(m_Geom.voxels(), T()); }
  const Geometry& GetGeometry() const { return m_Geom; }
  void SetGeometry(const Geometry& g) { m_Geom = g; }
  const std::array<int, 3>& GetSize() const { return m_Geom.size; }
  const std::array<double, 3>& GetSpacing() const { return m_Geom.spacing; }
  T* GetBufferPointer() { return m_Data.data(); }
  const T* GetBufferPointer() const { return m_Data.data(); }
  std::vector<T>& GetPixelContainer() { return m_Data; }
  size_t GetNumberOfPixels() const { return m_Data.size(); }
  T& GetPixel(int x, int y, int z) { return m_Data[(size_t)x + (size_t)m_Geom.size[0] * (y + (size_t)m_Geom.size[1] * z)]; }
  const T& GetPixel(int x, int y, int z) const { return m_Data[(size_t)x + (size_t)m_Geom.size[0] * (y + (size_t)m_Geom.size[1] * z)]; }

private:
  Geometry m_Geom;
  std::vector<T> m_Data;
};

// Multi-component image stored as SoA planes (component k = one contiguous volume), which
// is what the GPU writes and what every tool consumes (one file per component).
template <typename T>
class VectorImage {
public:
  typedef std::shared_ptr<VectorImage> Pointer;
  static Pointer New() { return std::make_shared<VectorImage>(); }
  void SetGeometry(const Geometry& g) { m_Geom = g; }
  const Geometry& GetGeometry() const { return m_Geom; }
  void SetNumberOfComponentsPerPixel(unsigned n) { m_Components = n; }
  unsigned GetNumberOfComponentsPerPixel() const { return m_Components; }
  void Allocate() { m_Data.assign(m_Geom.voxels() * m_Components, T()); }
  T* GetBufferPointer() { return m_Data.data(); }
  const T* GetComponentPointer(unsigned k) const { return m_Data.data() + (size_t)k * m_Geom.voxels(); }
  // itk::VectorIndexSelectionCastImageFilter equivalent
  typename Image<T>::Pointer ExtractComponent(unsigned k) const {
    if (k >= m_Components) throw std::out_of_range("component index");
    auto img = Image<T>::New();
    img->SetGeometry(m_Geom);
    img->GetPixelContainer().assign(GetComponentPointer(k), GetComponentPointer(k) + m_Geom.voxels());
    return img;
  }
  std::vector<T> GetPixel(int x, int y, int z) const {
    std::vector<T> p(m_Components);
    const size_t i = (size_t)x + (size_t)m_Geom.size[0] * (y + (size_t)m_Geom.size[1] * z);
    for (unsigned k = 0; k < m_Components; ++k) p[k] = m_Data[(size_t)k * m_Geom.voxels() + i];
    return p;
  }

private:
  Geometry m_Geom;
  unsigned m_Components = 0;
  std::vector<T> m_Data;
};

}  // namespace ife
#endif
