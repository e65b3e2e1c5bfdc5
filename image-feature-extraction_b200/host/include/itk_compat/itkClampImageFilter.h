// itk::ClampImageFilter (tools/ExtractFeatures.cxx:99-104 clamps the mask's labels to {0, 1}).
#ifndef IFE_B200_ITK_COMPAT_CLAMP_IMAGE_FILTER_H
#define IFE_B200_ITK_COMPAT_CLAMP_IMAGE_FILTER_H
#include <memory>

#include "itkImage.h"

namespace itk {
template <typename TInputImage, typename TOutputImage = TInputImage>
class ClampImageFilter {
public:
  typedef ClampImageFilter Self;
  typedef std::shared_ptr<Self> Pointer;
  typedef typename TOutputImage::PixelType OutputPixelType;
  static Pointer New() { return Pointer(new Self()); }
  void InPlaceOn() { m_InPlace = true; }
  void InPlaceOff() { m_InPlace = false; }
  void SetBounds(OutputPixelType lower, OutputPixelType upper) { m_Lower = lower; m_Upper = upper; }
  void SetInput(const TInputImage* image) { m_Input = image; m_Done = nullptr; }
  TOutputImage* GetOutput() { return m_Output.get(); }
  void Update() {
    if (!m_Input) throw ife::ExceptionObject(IFE_E_INVALID, "ClampImageFilter: input not set");
    m_Input->UpdateSource();
    if (m_Done == m_Input->GetBufferPointer()) return;
    const size_t n = m_Input->GetNumberOfPixels();
    m_Output->SetGeometry(m_Input->GetGeometry());
    m_Output->Allocate();
    const typename TInputImage::PixelType* in = m_Input->GetBufferPointer();
    OutputPixelType* out = m_Output->GetBufferPointer();
    for (size_t i = 0; i < n; ++i) {
      const typename TInputImage::PixelType v = in[i];
      out[i] = v < m_Lower ? m_Lower : (v > m_Upper ? m_Upper : static_cast<OutputPixelType>(v));
    }
    m_Done = m_Input->GetBufferPointer();
  }
private:
  ClampImageFilter() : m_Output(TOutputImage::New()) { m_Output->SetSource([this]() { this->Update(); }); }
  const TInputImage* m_Input = nullptr;
  const void* m_Done = nullptr;
  bool m_InPlace = false;
  OutputPixelType m_Lower = OutputPixelType(), m_Upper = OutputPixelType();
  typename TOutputImage::Pointer m_Output;
};
}  // namespace itk
#endif
