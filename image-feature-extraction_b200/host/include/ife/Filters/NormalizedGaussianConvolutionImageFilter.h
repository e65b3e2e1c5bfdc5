// Host-side mirror of itk::NormalizedGaussianConvolutionImageFilter
// (reference include/ife/Filters/NormalizedGaussianConvolutionImageFilter.h:54-113):
// SetInputImage / SetInputCertainty / SetSigma / GetSigma / Update / GetOutput; default
// sigma 1 (.hxx:18).  out = G(c*T)/G(c) on the GPU via ife_cuda_normalized_gaussian.
#ifndef IFE_B200_NORMALIZED_GAUSSIAN_CONVOLUTION_IMAGE_FILTER_H
#define IFE_B200_NORMALIZED_GAUSSIAN_CONVOLUTION_IMAGE_FILTER_H
#include <memory>

#include "ife/Context.h"
#include "ife/Image.h"

namespace ife {

template <typename TImage = Image<float> >
class NormalizedGaussianConvolutionImageFilter {
public:
  typedef NormalizedGaussianConvolutionImageFilter Self;
  typedef std::shared_ptr<Self> Pointer;
  typedef TImage ImageType;
  typedef double ScalarRealType;
  static Pointer New() { return Pointer(new Self()); }

  void SetInputImage(const ImageType* image) { m_Image = image; }
  void SetInputCertainty(const ImageType* certainty) { m_Certainty = certainty; }
  void SetSigma(ScalarRealType s) { m_Sigma = s; }
  ScalarRealType GetSigma() const { return m_Sigma; }
  // extension used by the MaskedNormalizedConvolution tool's -m flag (itk::MaskImageFilter
  // with the certainty as mask, tools/MaskedNormalizedConvolution.cxx:156-159), fused on the GPU
  void SetMaskOutput(bool m) { m_MaskOutput = m; }

  void Update() {
    if (!m_Image || !m_Certainty) throw ExceptionObject(IFE_E_INVALID, "NormalizedGaussianConvolutionImageFilter: inputs not set");
    m_Image->UpdateSource();
    m_Certainty->UpdateSource();
    const Geometry& g = m_Image->GetGeometry();
    if (m_Certainty->GetGeometry().size != g.size)
      throw ExceptionObject(IFE_E_INVALID, "NormalizedGaussianConvolutionImageFilter: image and certainty sizes differ");
    m_Output->SetGeometry(g);
    m_Output->Allocate();
    CudaContext& c = CudaContext::Instance();
    c.Check(ife_cuda_normalized_gaussian(c.Handle(), m_Image->GetBufferPointer(), m_Certainty->GetBufferPointer(),
                                         nullptr, m_Output->GetBufferPointer(), g.size.data(), g.spacing.data(),
                                         m_Sigma, m_MaskOutput ? 1 : 0, IFE_MEM_HOST));
  }
  ImageType* GetOutput() { return m_Output.get(); }

private:
  NormalizedGaussianConvolutionImageFilter() : m_Output(ImageType::New()) { m_Output->SetSource([this]() { this->Update(); }); }
  const ImageType* m_Image = nullptr;
  const ImageType* m_Certainty = nullptr;
  ScalarRealType m_Sigma = 1.0;
  bool m_MaskOutput = false;
  typename ImageType::Pointer m_Output;
};

}  // namespace ife

namespace itk {
using ife::NormalizedGaussianConvolutionImageFilter;
}
#endif
