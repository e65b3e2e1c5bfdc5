// itk::VectorIndexSelectionCastImageFilter: component `index` of a VectorImage as an image.
// The facades store components as planes, so the output is a VIEW of the input's storage.
#ifndef IFE_B200_ITK_COMPAT_VECTOR_INDEX_SELECTION_H
#define IFE_B200_ITK_COMPAT_VECTOR_INDEX_SELECTION_H
#include <memory>

#include "itkImage.h"

namespace itk {
template <typename TInputImage, typename TOutputImage>
class VectorIndexSelectionCastImageFilter {
public:
  typedef VectorIndexSelectionCastImageFilter Self;
  typedef std::shared_ptr<Self> Pointer;
  static Pointer New() { return Pointer(new Self()); }
  void SetInput(const TInputImage* image) { m_Input = image; }
  void SetIndex(unsigned int i) { m_Index = i; }
  unsigned int GetIndex() const { return m_Index; }
  TOutputImage* GetOutput() { return m_Output.get(); }
  void Update() {
    if (!m_Input) throw ife::ExceptionObject(IFE_E_INVALID, "VectorIndexSelectionCastImageFilter: input not set");
    m_Input->UpdateSource();
    if (m_Index >= m_Input->GetNumberOfComponentsPerPixel())
      throw ife::ExceptionObject(IFE_E_INVALID, "VectorIndexSelectionCastImageFilter: index out of range");
    const size_t n = m_Input->GetGeometry().voxels();
    m_Output->SetView(m_Input->GetGeometry(),
                      const_cast<typename TOutputImage::PixelType*>(m_Input->GetComponentPointer(m_Index)), n,
                      m_Input->keepalive());
  }
private:
  VectorIndexSelectionCastImageFilter() : m_Output(TOutputImage::New()) {
    m_Output->SetSource([this]() { this->Update(); });
  }
  const TInputImage* m_Input = nullptr;
  unsigned int m_Index = 0;
  typename TOutputImage::Pointer m_Output;
};
}  // namespace itk
#endif
