// FiniteDifference_HessianFeatures -i/--image -m/--mask -o/--outdir [-p/--prefix hessian_]
//                                  [-s/--sigma 0] [-b/--reference-dy-bug 0]
// Flags, semantics and output names of the reference tool
// (tools/FiniteDifference_HessianFeatures.cxx:45-82,127-229,253-264): central-difference
// Hessian of the image, per-voxel eigen features where mask != 0 and zeros elsewhere, six
// outputs <outdir>/<prefix>{eig1,eig2,eig3,LoG,Curvature,Frobenius}.nii.gz.
// Two additions: -s smooths first with the library Gaussian (BASELINE.json configs[0]; the
// default 0 is the tool as shipped), and -b 1 reproduces the shipped source's
// `dyFilter->SetDirection(0)` (:153-156), which turns its Dyz into Dz(Dx).  The reference
// tool is excluded from its own build (tools/CMakeLists.txt:32) and does not compile
// (missing Eigenvalues.h), so the default here is the Hessian of Hessian3DImageFilter.
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "ife/Filters/Hessian3DImageFilter.h"
#include "ife/IO/NiftiIO.h"
#include "ife/Util/CmdLine.h"
#include "ife/Util/Path.h"

const std::string VERSION("0.1");
const std::string OUT_FILE_TYPE(".nii.gz");

int main(int argc, char* argv[]) {
  ife::CmdLine cmd("Calculate Hessian based features.", VERSION);
  cmd.add("i", "image", "Path to image.", true, "", "path");
  cmd.add("m", "mask", "Path to mask. Must match image dimensions.", true, "", "path");
  cmd.add("o", "outdir", "Path to output directory", true, "", "path");
  cmd.add("p", "prefix", "Prefix to use for output filenames", false, "hessian_", "string");
  cmd.add("s", "sigma", "Smooth with a Gaussian of this scale first (0 = no smoothing)", false, "0", "double");
  cmd.add("b", "reference-dy-bug", "Reproduce the shipped tool's dy direction (Dyz := Dz(Dx))", false, "0", "boolean");
  int rc;
  if (!cmd.parse(argc, argv, &rc)) return rc;
  const std::string imagePath(cmd.value("image")), maskPath(cmd.value("mask"));
  const std::string outDirPath(cmd.value("outdir")), prefix(cmd.value("prefix"));
  double sigma = 0;
  bool dyBug = false;
  if (!ife::CmdLine::convert(cmd.value("sigma"), &sigma)) { cmd.error("Couldn't read argument value", "-s", &rc); return rc; }
  if (!ife::CmdLine::to_bool(cmd.value("reference-dy-bug"), &dyBug)) { cmd.error("Couldn't read argument value", "-b", &rc); return rc; }

  const std::string baseFileName = ife::Path::join(outDirPath, prefix);
  const std::vector<std::string> featureNames{"eig1", "eig2", "eig3", "LoG", "Curvature", "Frobenius"};
  try {
    auto both = ife::nifti::ReadPair<float, unsigned char>(imagePath, maskPath);   // the two files are inflated concurrently
    ife::Image<float>::Pointer image = both.first;
    ife::Image<unsigned char>::Pointer mask = both.second;
    if (mask->GetSize() != image->GetSize()) throw std::runtime_error("mask and image dimensions differ");
    auto hessianFilter = ife::HessianEigenFeaturesImageFilter<>::New();
    hessianFilter->SetInput(image.get());
    hessianFilter->SetMask(mask.get());
    hessianFilter->SetSigma(sigma);
    hessianFilter->SetReproduceToolDirectionBug(dyBug);
    hessianFilter->Update();
    for (unsigned int i = 0; i < hessianFilter->GetOutput()->GetNumberOfComponentsPerPixel(); ++i) {
      const std::string outFile = baseFileName + featureNames.at(i) + OUT_FILE_TYPE;
      ife::nifti::Write(outFile, image->GetGeometry(), hessianFilter->GetOutput()->GetComponentPointer(i));
    }
  } catch (std::exception& e) {
    std::cerr << "Failed to process." << std::endl
              << "Image: " << imagePath << std::endl
              << "Mask: " << maskPath << std::endl
              << "Base file name: " << baseFileName << std::endl
              << "ExceptionObject: " << e.what() << std::endl;
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}
