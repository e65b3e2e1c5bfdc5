// Host-side mirror of itk::ImageToEmphysemaFeaturesFilter
// (reference include/ife/Filters/ImageToEmphysemaFeaturesFilter.h:21-145): same class name,
// New / SetInputImage / SetInputMask / SetSigma / GetSigma / Update /
// UpdateLargestPossibleRegion / GetOutput / numFeatures, computed on the GPU through
// ife_cuda_emphysema_features.  Output: 8-component image
// [GaussianBlur, GradientMagnitude, Eigenvalue1..3, LoG, GaussianCurvature, FrobeniusNorm].
#ifndef IFE_B200_IMAGE_TO_EMPHYSEMA_FEATURES_FILTER_H
#define IFE_B200_IMAGE_TO_EMPHYSEMA_FEATURES_FILTER_H
#include <memory>
#include <vector>

#include "ife/Context.h"
#include "ife/Image.h"

namespace ife {

template <typename TInputImage = Image<float>, typename TInputMask = Image<unsigned char>,
          typename TOutputImage = VectorImage<float> >
class ImageToEmphysemaFeaturesFilter {
public:
  typedef ImageToEmphysemaFeaturesFilter Self;
  typedef std::shared_ptr<Self> Pointer;
  typedef TInputImage InputImageType;
  typedef TInputMask InputMaskType;
  typedef TOutputImage OutputImageType;
  typedef float ScalarRealType;
  static const size_t numFeatures = 8;

  static Pointer New() { return Pointer(new Self()); }

  void SetInputImage(const InputImageType* image) { m_Image = image; m_Modified = true; }
  void SetInputMask(const InputMaskType* mask) { m_Mask = mask; m_Modified = true; }
  void SetSigma(ScalarRealType sigma) { if (sigma != m_Sigma) { m_Sigma = sigma; m_Modified = true; } }
  ScalarRealType GetSigma() const { return m_Sigma; }

  // All scales in one GPU call (one H2D of the inputs); GetOutput(i) then selects scale i.
  void SetSigmas(const std::vector<double>& sigmas) { m_Sigmas = sigmas; m_Modified = true; }

  void Update() {
    if (!m_Modified && m_Outputs.size()) return;
    if (!m_Image || !m_Mask) throw ExceptionObject(IFE_E_INVALID, "ImageToEmphysemaFeaturesFilter: inputs not set");
    const Geometry& g = m_Image->GetGeometry();
    if (m_Mask->GetGeometry().size != g.size)
      throw ExceptionObject(IFE_E_INVALID, "ImageToEmphysemaFeaturesFilter: image and mask sizes differ");
    std::vector<double> sigmas = m_Sigmas.empty() ? std::vector<double>(1, (double)m_Sigma) : m_Sigmas;
    const size_t n = g.voxels();
    std::vector<float> buf(sigmas.size() * numFeatures * n);
    CudaContext& c = CudaContext::Instance();
    c.Check(ife_cuda_emphysema_features(c.Handle(), m_Image->GetBufferPointer(), m_Mask->GetBufferPointer(),
                                        buf.data(), g.size.data(), g.spacing.data(), sigmas.data(),
                                        (int)sigmas.size(), IFE_MEM_HOST));
    m_Outputs.clear();
    for (size_t s = 0; s < sigmas.size(); ++s) {
      auto out = OutputImageType::New();
      out->SetGeometry(g);
      out->SetNumberOfComponentsPerPixel(numFeatures);
      out->Allocate();
      std::copy(buf.begin() + s * numFeatures * n, buf.begin() + (s + 1) * numFeatures * n, out->GetBufferPointer());
      m_Outputs.push_back(out);
    }
    m_Modified = false;
  }
  void UpdateLargestPossibleRegion() { Update(); }
  OutputImageType* GetOutput(size_t scale = 0) { return m_Outputs.at(scale).get(); }

private:
  ImageToEmphysemaFeaturesFilter() {}
  const InputImageType* m_Image = nullptr;
  const InputMaskType* m_Mask = nullptr;
  ScalarRealType m_Sigma = 1.0f;   // reference default (ImageToEmphysemaFeaturesFilter.hxx:18)
  std::vector<double> m_Sigmas;
  std::vector<typename OutputImageType::Pointer> m_Outputs;
  bool m_Modified = true;
};

}  // namespace ife
#endif
