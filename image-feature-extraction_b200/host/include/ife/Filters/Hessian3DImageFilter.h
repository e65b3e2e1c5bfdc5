// Host-side mirror of itk::Hessian3DImageFilter
// (reference include/ife/Filters/Hessian3DImageFilter.h:23-83, wiring .hxx:11-60): SetInput /
// Update / GetOutput, output = the 6-component Hessian [Dxx, Dxy, Dxz, Dyy, Dyz, Dzz]
// (.hxx:53-59) of the image as given, computed on the GPU by ife_cuda_hessian.
//
// HessianEigenFeaturesImageFilter is the fused form the FiniteDifference_HessianFeatures tool
// uses: Hessian -> EigenvalueFeaturesFunctor -> mask in one kernel, the Hessian never
// written (tools/FiniteDifference_HessianFeatures.cxx:127-229), via
// ife_cuda_hessian_eigen_features.
#ifndef IFE_B200_HESSIAN3D_IMAGE_FILTER_H
#define IFE_B200_HESSIAN3D_IMAGE_FILTER_H
#include <memory>

#include "ife/Context.h"
#include "ife/Image.h"

namespace ife {

template <typename TInputImage = Image<float>, typename TOutputImage = VectorImage<float> >
class Hessian3DImageFilter {
public:
  typedef Hessian3DImageFilter Self;
  typedef std::shared_ptr<Self> Pointer;
  typedef TInputImage InputImageType;
  typedef TOutputImage OutputImageType;
  static const unsigned int NumberOfComponents = 6;
  static Pointer New() { return Pointer(new Self()); }

  void SetInput(const InputImageType* image) { m_Image = image; }
  void Update() {
    if (!m_Image) throw ExceptionObject(IFE_E_INVALID, "Hessian3DImageFilter: input not set");
    m_Image->UpdateSource();
    const Geometry& g = m_Image->GetGeometry();
    m_Output->SetGeometry(g);
    m_Output->SetNumberOfComponentsPerPixel(NumberOfComponents);
    m_Output->Allocate();
    CudaContext& c = CudaContext::Instance();
    c.Check(ife_cuda_hessian(c.Handle(), m_Image->GetBufferPointer(), m_Output->GetBufferPointer(), g.size.data(),
                             g.spacing.data(), IFE_MEM_HOST));
  }
  void UpdateLargestPossibleRegion() { Update(); }
  OutputImageType* GetOutput() { return m_Output.get(); }

private:
  Hessian3DImageFilter() : m_Output(OutputImageType::New()) { m_Output->SetSource([this]() { this->Update(); }); }
  const InputImageType* m_Image = nullptr;
  typename OutputImageType::Pointer m_Output;
};

template <typename TInputImage = Image<float>, typename TOutputImage = VectorImage<float> >
class HessianEigenFeaturesImageFilter {
public:
  typedef HessianEigenFeaturesImageFilter Self;
  typedef std::shared_ptr<Self> Pointer;
  typedef TInputImage InputImageType;
  typedef TOutputImage OutputImageType;
  static Pointer New() { return Pointer(new Self()); }

  void SetInput(const InputImageType* image) { m_Image = image; }
  void SetMask(const Image<unsigned char>* mask) { m_Mask = mask; }
  void SetSigma(double sigma) { m_Sigma = sigma; }       // <= 0: no smoothing (the tool as shipped)
  void SetReproduceToolDirectionBug(bool b) { m_DyBug = b; }

  void Update() {
    if (!m_Image) throw ExceptionObject(IFE_E_INVALID, "HessianEigenFeaturesImageFilter: input not set");
    m_Image->UpdateSource();
    if (m_Mask) m_Mask->UpdateSource();
    const Geometry& g = m_Image->GetGeometry();
    m_Output->SetGeometry(g);
    m_Output->SetNumberOfComponentsPerPixel(IFE_NUM_EIGEN_FEATURES);
    m_Output->Allocate();
    CudaContext& c = CudaContext::Instance();
    c.Check(ife_cuda_hessian_eigen_features(c.Handle(), m_Image->GetBufferPointer(),
                                            m_Mask ? m_Mask->GetBufferPointer() : nullptr,
                                            m_Output->GetBufferPointer(), g.size.data(), g.spacing.data(),
                                            m_Sigma, m_DyBug ? IFE_FDHF_TOOL_DY_BUG : 0, IFE_MEM_HOST));
  }
  OutputImageType* GetOutput() { return m_Output.get(); }

private:
  HessianEigenFeaturesImageFilter() : m_Output(OutputImageType::New()) { m_Output->SetSource([this]() { this->Update(); }); }
  const InputImageType* m_Image = nullptr;
  const Image<unsigned char>* m_Mask = nullptr;
  double m_Sigma = 0.0;
  bool m_DyBug = false;
  typename OutputImageType::Pointer m_Output;
};

}  // namespace ife

namespace itk {
using ife::Hessian3DImageFilter;
}
#endif
