// ExtractFeatures -i/--image -m/--mask -o/--out -s/--scale ...
// Same flags, semantics and output names as the reference tool (tools/ExtractFeatures.cxx):
// float image, unsigned-char mask clamped to {0,1} (:99-104), for every scale the 8 features
// of ImageToEmphysemaFeaturesFilter written as
//   <out>_scale_<std::to_string(float sigma)><FeatureName>.nii.gz            (:132-139)
// All scales are computed in one GPU call (one upload of the scan).
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "itkImageFileReader.h"
#include "itkImageFileWriter.h"
#include "itkVectorIndexSelectionCastImageFilter.h"
#include "itkClampImageFilter.h"

#include "ife/Filters/ImageToEmphysemaFeaturesFilter.h"
#include "ife/Util/CmdLine.h"

const std::string VERSION("0.1");
const std::string OUT_FILE_TYPE(".nii.gz");

int main(int argc, char* argv[]) {
  ife::CmdLine cmd("Create a bag of instances samples from an image.", VERSION);
  cmd.add("i", "image", "Path to image.", true, "", "path");
  cmd.add("m", "mask", "Path to mask.", true, "", "path");
  cmd.add("o", "out", "Base output path", true, "", "path");
  cmd.add("s", "scale", "Scales for the Gauss applicability function", true, "", "double", true);
  int rc;
  if (!cmd.parse(argc, argv, &rc)) return rc;
  const std::string imagePath(cmd.value("image")), maskPath(cmd.value("mask")), outBasePath(cmd.value("out"));
  std::vector<float> scales;
  for (const std::string& s : cmd.values("scale")) {
    float v;
    if (!ife::CmdLine::convert(s, &v)) { cmd.error("Couldn't read argument value from string '" + s + "'", "-s", &rc); return rc; }
    scales.push_back(v);
  }

  // The reference's pipeline (tools/ExtractFeatures.cxx:76-154), spelled with the same itk:: names:
  // readers -> clamp(0, 1) on the mask -> feature filter -> component selection -> writer, pulled
  // by writer->Update().  One difference: the scales are handed over together, so the scan is
  // uploaded once and the download of scale s overlaps the kernels of scale s+1.
  typedef float PixelType;
  typedef unsigned char MaskPixelType;
  const unsigned int Dimension = 3;
  typedef itk::Image<PixelType, Dimension> ImageType;
  typedef itk::Image<MaskPixelType, Dimension> MaskType;
  typedef itk::VectorImage<PixelType, Dimension> VectorImageType;

  typedef itk::ImageFileReader<ImageType> ReaderType;
  ReaderType::Pointer reader = ReaderType::New();
  reader->SetFileName(imagePath);
  typedef itk::ImageFileReader<MaskType> MaskReaderType;
  MaskReaderType::Pointer maskReader = MaskReaderType::New();
  maskReader->SetFileName(maskPath);

  typedef itk::ClampImageFilter<MaskType, MaskType> ClampFilterType;
  ClampFilterType::Pointer clampFilter = ClampFilterType::New();
  clampFilter->InPlaceOn();
  clampFilter->SetBounds(0, 1);
  clampFilter->SetInput(maskReader->GetOutput());

  typedef itk::ImageToEmphysemaFeaturesFilter<ImageType, MaskType, VectorImageType> FeatureFilterType;
  FeatureFilterType::Pointer featureFilter = FeatureFilterType::New();
  featureFilter->SetInputImage(reader->GetOutput());
  featureFilter->SetInputMask(clampFilter->GetOutput());
  featureFilter->SetSigmas(std::vector<double>(scales.begin(), scales.end()));

  typedef itk::VectorIndexSelectionCastImageFilter<VectorImageType, ImageType> IndexSelectionType;
  IndexSelectionType::Pointer indexSelectionFilter = IndexSelectionType::New();
  typedef itk::ImageFileWriter<ImageType> WriterType;
  WriterType::Pointer writer = WriterType::New();
  writer->SetInput(indexSelectionFilter->GetOutput());

  const std::vector<std::string> featureNames{"GaussianBlur", "GradientMagnitude", "Eigenvalue1", "Eigenvalue2",
                                              "Eigenvalue3", "LaplacianOfGaussian", "GaussianCurvature", "FrobeniusNorm"};
  for (size_t s = 0; s < scales.size(); ++s) {
    indexSelectionFilter->SetInput(featureFilter->GetOutput(s));
    for (unsigned int i = 0; i < featureNames.size(); ++i) {
      indexSelectionFilter->SetIndex(i);
      const std::string outPath = outBasePath + "_scale_" + std::to_string(scales[s]) + featureNames[i] + OUT_FILE_TYPE;
      writer->SetFileName(outPath);
      try {
        featureFilter->UpdateLargestPossibleRegion();
        writer->Update();
      } catch (itk::ExceptionObject& e) {
        std::cerr << "Failed to process." << std::endl
                  << "Image: " << imagePath << std::endl
                  << "Mask: " << maskPath << std::endl
                  << "Out: " << outPath << std::endl
                  << "ExceptionObject: " << e << std::endl;
        return EXIT_FAILURE;
      }
    }
  }
  if (std::getenv("IFE_TIMING"))
    std::cerr << "[timing] writing " << scales.size() * featureNames.size() << " " << OUT_FILE_TYPE << " files: "
              << ife::nifti::write_seconds() << " s; reading: " << ife::nifti::read_seconds()
              << " s; page-locked allocation: " << ife::alloc_seconds() << " s; ife_cuda_emphysema_features(IFE_MEM_HOST, page-locked buffers, " << scales.size()
              << " scales): " << featureFilter->GetLastCallSeconds() << " s" << std::endl;
  return EXIT_SUCCESS;
}
