// MaskedNormalizedConvolution -i/--image -c/--certainty -s/--scale ... -o/--outdir
//                             [-p/--prefix normconv_] [-m/--maskoutput bool]
// Same flags, semantics and output names as the reference tool
// (tools/MaskedNormalizedConvolution.cxx): image and certainty read as float (:129-139),
// out = G(cT)/G(c) per scale, optionally masked by the certainty (:156-159), written as
//   <outdir>/<prefix>scale_<std::to_string(double sigma)>.nii.gz              (:179-188)
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "ife/Filters/NormalizedGaussianConvolutionImageFilter.h"
#include "ife/IO/NiftiIO.h"
#include "ife/Util/CmdLine.h"
#include "ife/Util/Path.h"

const std::string VERSION("0.1");
const std::string OUT_FILE_TYPE(".nii.gz");

int main(int argc, char* argv[]) {
  ife::CmdLine cmd("Perform normalized convolution.", VERSION);
  cmd.add("i", "image", "Path to image (T)", true, "", "path");
  cmd.add("c", "certainty", "Path to certainty mask (c). Must match image dimensions.", true, "", "path");
  cmd.add("s", "scale", "Scales for the Gauss applicability function", true, "", "double", true);
  cmd.add("o", "outdir", "Path to output directory", true, "", "path");
  cmd.add("p", "prefix", "Prefix to use for output filenames", false, "normconv_", "string");
  cmd.add("m", "maskoutput", "Mask the output after convolution.", false, "0", "boolean");
  int rc;
  if (!cmd.parse(argc, argv, &rc)) return rc;
  const std::string imagePath(cmd.value("image")), certaintyPath(cmd.value("certainty"));
  const std::string outDirPath(cmd.value("outdir")), prefix(cmd.value("prefix"));
  std::vector<double> scales;
  for (const std::string& s : cmd.values("scale")) {
    double v;
    if (!ife::CmdLine::convert(s, &v)) { cmd.error("Couldn't read argument value from string '" + s + "'", "-s", &rc); return rc; }
    scales.push_back(v);
  }
  bool maskOutput = false;
  if (!ife::CmdLine::to_bool(cmd.value("maskoutput"), &maskOutput)) {
    cmd.error("Couldn't read argument value from string '" + cmd.value("maskoutput") + "'", "-m", &rc);
    return rc;
  }

  typedef ife::Image<float> ImageType;
  const std::string baseFileName = ife::Path::join(outDirPath, prefix);
  double scale = 0;
  try {
    auto both = ife::nifti::ReadPair<float, float>(imagePath, certaintyPath);   // the two files are inflated concurrently
    ImageType::Pointer image = both.first;
    ImageType::Pointer certainty = both.second;
    auto normConvFilter = ife::NormalizedGaussianConvolutionImageFilter<>::New();
    normConvFilter->SetInputImage(image.get());
    normConvFilter->SetInputCertainty(certainty.get());
    normConvFilter->SetMaskOutput(maskOutput);
    for (double s : scales) {
      scale = s;
      std::cout << "Processing scale " << scale << std::endl;
      normConvFilter->SetSigma(scale);
      normConvFilter->Update();
      const std::string outFile = baseFileName + "scale_" + std::to_string(scale) + OUT_FILE_TYPE;
      ife::nifti::Write(outFile, *normConvFilter->GetOutput());
    }
  } catch (std::exception& e) {
    std::cerr << "Failed to process." << std::endl
              << "Image: " << imagePath << std::endl
              << "Certainty: " << certaintyPath << std::endl
              << "Scale: " << scale << std::endl
              << "Base file name: " << baseFileName << std::endl
              << "ExceptionObject: " << e.what() << std::endl;
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}
