// Host-side mirror of itk::ImageToEmphysemaFeaturesFilter
// (reference include/ife/Filters/ImageToEmphysemaFeaturesFilter.h:21-145): same class name,
// template parameters <TInputImage, TInputMask, TOutputImage>, New / SetInputImage /
// SetInputMask / SetSigma / GetSigma / Update / UpdateLargestPossibleRegion / GetOutput /
// numFeatures, computed on the GPU through ife_cuda_emphysema_features.  Output: 8-component
// image [GaussianBlur, GradientMagnitude, Eigenvalue1..3, LoG, GaussianCurvature, FrobeniusNorm].
//
// The result of every scale lives in ONE page-locked buffer that the C ABI fills directly
// (no value-initialised staging vector, no second copy); GetOutput(s) is a view into it.
// GetOutput() exists before the first Update(), as in ITK, so that a pipeline can be wired
// first and pulled later (tools/ExtractFeatures.cxx:106-154).
#ifndef IFE_B200_IMAGE_TO_EMPHYSEMA_FEATURES_FILTER_H
#define IFE_B200_IMAGE_TO_EMPHYSEMA_FEATURES_FILTER_H
#include <chrono>
#include <memory>
#include <future>
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
  void SetSigma(ScalarRealType sigma) {
    if (sigma != m_Sigma || !m_Sigmas.empty()) { m_Sigma = sigma; m_Sigmas.clear(); m_Modified = true; }
  }
  ScalarRealType GetSigma() const { return m_Sigma; }

  // All scales in one GPU call (one H2D of the inputs); GetOutput(i) then selects scale i.
  void SetSigmas(const std::vector<double>& sigmas) {
    m_Sigmas = sigmas;
    m_Modified = true;
    while (m_Outputs.size() < sigmas.size()) m_Outputs.push_back(NewOutput());
  }

  void Update() {
    if (!m_Image || !m_Mask) throw ExceptionObject(IFE_E_INVALID, "ImageToEmphysemaFeaturesFilter: inputs not set");
    {   // the two upstream pipelines (typically two file readers) run side by side
      const auto* mask = m_Mask;
      auto other = std::async(std::launch::async, [mask]() { mask->UpdateSource(); });
      m_Image->UpdateSource();
      other.get();
    }
    if (!m_Modified && m_Image->GetBufferPointer() == m_LastImage && m_Mask->GetBufferPointer() == m_LastMask) return;
    const Geometry& g = m_Image->GetGeometry();
    if (m_Mask->GetGeometry().size != g.size)
      throw ExceptionObject(IFE_E_INVALID, "ImageToEmphysemaFeaturesFilter: image and mask sizes differ");
    std::vector<double> sigmas = m_Sigmas.empty() ? std::vector<double>(1, (double)m_Sigma) : m_Sigmas;
    const size_t n = g.voxels();
    if (!m_Store) m_Store = std::make_shared<PixelStore<float> >();
    m_Store->data.resize(sigmas.size() * numFeatures * n);   // page-locked, uninitialised
    CudaContext& c = CudaContext::Instance();
    ife_cuda_ctx* h = c.Handle();
    const auto t0 = std::chrono::steady_clock::now();
    c.Check(ife_cuda_emphysema_features(h, m_Image->GetBufferPointer(), m_Mask->GetBufferPointer(),
                                        m_Store->data.data(), g.size.data(), g.spacing.data(), sigmas.data(),
                                        (int)sigmas.size(), IFE_MEM_HOST));
    m_LastCallSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (size_t s = 0; s < sigmas.size(); ++s)
      m_Outputs[s]->SetView(g, (unsigned)numFeatures, m_Store->data.data() + s * numFeatures * n, m_Store);
    m_LastImage = m_Image->GetBufferPointer();
    m_LastMask = m_Mask->GetBufferPointer();
    m_Modified = false;
  }
  void UpdateLargestPossibleRegion() { Update(); }
  OutputImageType* GetOutput(size_t scale = 0) { return m_Outputs.at(scale).get(); }
  // wall time of the last ife_cuda_emphysema_features call (uploads, kernels, downloads)
  double GetLastCallSeconds() const { return m_LastCallSeconds; }

private:
  ImageToEmphysemaFeaturesFilter() { m_Outputs.push_back(NewOutput()); }
  typename OutputImageType::Pointer NewOutput() {
    auto out = OutputImageType::New();
    out->SetNumberOfComponentsPerPixel((unsigned)numFeatures);
    out->SetSource([this]() { this->Update(); });
    return out;
  }
  const InputImageType* m_Image = nullptr;
  const InputMaskType* m_Mask = nullptr;
  const void* m_LastImage = nullptr;
  const void* m_LastMask = nullptr;
  ScalarRealType m_Sigma = 1.0f;   // reference default (ImageToEmphysemaFeaturesFilter.hxx:18)
  std::vector<double> m_Sigmas;
  std::shared_ptr<PixelStore<float> > m_Store;
  std::vector<typename OutputImageType::Pointer> m_Outputs;
  bool m_Modified = true;
  double m_LastCallSeconds = 0.0;
};

}  // namespace ife

// the reference spells these classes itk:: (its headers open namespace itk)
namespace itk {
using ife::ImageToEmphysemaFeaturesFilter;
}
#endif
