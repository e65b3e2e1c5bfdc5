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

#include "ife/Filters/ImageToEmphysemaFeaturesFilter.h"
#include "ife/IO/NiftiIO.h"
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

  typedef ife::Image<float> ImageType;
  typedef ife::Image<unsigned char> MaskType;
  const std::vector<std::string> featureNames{"GaussianBlur", "GradientMagnitude", "Eigenvalue1", "Eigenvalue2",
                                              "Eigenvalue3", "LaplacianOfGaussian", "GaussianCurvature", "FrobeniusNorm"};
  std::string outPath;
  try {
    ImageType::Pointer image = ife::nifti::Read<float>(imagePath);
    MaskType::Pointer mask = ife::nifti::Read<unsigned char>(maskPath);
    for (unsigned char& v : mask->GetPixelContainer()) v = v > 1 ? 1 : v;  // ClampImageFilter(0, 1)

    auto featureFilter = ife::ImageToEmphysemaFeaturesFilter<>::New();
    featureFilter->SetInputImage(image.get());
    featureFilter->SetInputMask(mask.get());
    featureFilter->SetSigmas(std::vector<double>(scales.begin(), scales.end()));
    featureFilter->UpdateLargestPossibleRegion();

    for (size_t s = 0; s < scales.size(); ++s) {
      for (unsigned int i = 0; i < featureNames.size(); ++i) {
        outPath = outBasePath + "_scale_" + std::to_string(scales[s]) + featureNames[i] + OUT_FILE_TYPE;
        ife::nifti::Write(outPath, image->GetGeometry(), featureFilter->GetOutput(s)->GetComponentPointer(i));
      }
    }
  } catch (std::exception& e) {
    std::cerr << "Failed to process." << std::endl
              << "Image: " << imagePath << std::endl
              << "Mask: " << maskPath << std::endl
              << "Out: " << outPath << std::endl
              << "ExceptionObject: " << e.what() << std::endl;
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}
