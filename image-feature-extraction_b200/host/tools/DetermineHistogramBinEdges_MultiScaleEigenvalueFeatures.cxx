// DetermineHistogramBinEdges_MultiScaleEigenvalueFeatures -i pairlist -o outfile -b bins
//        -S samples -s scale ... -f foreground ... [--seed n]
// Flags, formats and semantics of the reference tool
// (tools/DetermineHistogramBinEdges_MultiScaleEigenvalueFeatures.cxx): for every
// "image,mask" line of the pair list (src/IO/IO.cxx:20-41) the 8 features at every scale;
// samples are taken at every voxel whose mask value is one of the foreground values
// (-S 0) or at -S random such voxels per scale (:171-264); per (scale, feature) the samples
// are sorted and equal-frequency edges determined (:282-296); the output is the
// histogram-spec file MakeBag reads (two '#' header lines, one row of bins-1 edges each).
// Features are computed on the GPU and only the sampled values leave it (compaction sink); the
// sort is the library's own device radix sort.
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <random>
#include <string>
#include <vector>

#include "ife/Context.h"
#include "ife/IO/NiftiIO.h"
#include "ife/Statistics/DetermineEdgesForEqualizedHistogram.h"
#include "ife/Util/CmdLine.h"

const std::string VERSION("0.1");

static std::string trim(const std::string& s, const std::string& chars = " \r\n\t") {
  const size_t a = s.find_first_not_of(chars), b = s.find_last_not_of(chars);
  return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}

int main(int argc, char* argv[]) {
  ife::CmdLine cmd("Find bin edges for equalized histograms of multi scale eigenvalue features.", VERSION);
  cmd.add("i", "infile", "Path to image/mask list.", true, "", "path");
  cmd.add("o", "outfile", "Path to output file", true, "", "path");
  cmd.add("b", "bins", "Number of bins to use", true, "41", "unsigned int");
  cmd.add("S", "samples", "Number of samples to use from each (0 = all)", true, "0", "unsigned int");
  cmd.add("s", "scale", "Scales for the Gauss applicability function", true, "", "double", true);
  cmd.add("f", "foreground", "Voxel value of foreground in mask", true, "", "unsigned int", true);
  cmd.add("e", "seed", "Seed for the random sampling (default: non-deterministic)", false, "", "integer");
  int rc;
  if (!cmd.parse(argc, argv, &rc)) return rc;
  const std::string infilePath(cmd.value("infile")), outfilePath(cmd.value("outfile"));
  unsigned int nBins = 0, nSamples = 0;
  std::vector<float> scales;
  std::vector<unsigned int> foreground;
  bool ok = ife::CmdLine::convert(cmd.value("bins"), &nBins) && ife::CmdLine::convert(cmd.value("samples"), &nSamples);
  for (const std::string& s : cmd.values("scale")) { float v; ok = ok && ife::CmdLine::convert(s, &v); scales.push_back(v); }
  for (const std::string& s : cmd.values("foreground")) { unsigned v; ok = ok && ife::CmdLine::convert(s, &v); foreground.push_back(v); }
  if (!ok) { cmd.error("Couldn't read a numeric argument value", "", &rc); return rc; }

  std::vector<std::pair<std::string, std::string> > pairs;
  {
    std::ifstream is(infilePath);
    std::string line;
    bool bad = !is.good();
    while (!bad && std::getline(is, line)) {
      if (line.empty()) continue;
      const size_t pos = line.find(',');
      if (pos == std::string::npos) { bad = true; break; }
      pairs.emplace_back(trim(line.substr(0, pos), " "), trim(line.substr(pos + 1)));
    }
    if (bad) { std::cerr << "Could not read image/mask list" << std::endl; return EXIT_FAILURE; }
  }

  const size_t numFeatures = 8;
  std::vector<std::vector<float> > samples(scales.size() * numFeatures);
  std::mt19937_64 gen;
  if (cmd.value("seed").empty()) gen.seed(std::random_device{}());
  else { unsigned long long s = 0; ife::CmdLine::convert(cmd.value("seed"), &s); gen.seed(s); }

  for (const auto& pr : pairs) {
    std::cout << "Processing " << std::endl << "Image: '" << pr.first << "'" << std::endl << "Mask: '" << pr.second << "'" << std::endl;
    try {
      auto both = ife::nifti::ReadPair<float, unsigned short>(pr.first, pr.second);   // the two files are inflated concurrently
      auto image = both.first;
      auto mask16 = both.second;
      if (mask16->GetSize() != image->GetSize()) throw std::runtime_error("mask and image dimensions differ");
      const size_t n = image->GetNumberOfPixels();
      auto mask = ife::Image<unsigned char>::New();
      mask->SetGeometry(image->GetGeometry());
      mask->Allocate();
      std::vector<unsigned char> fg(n);
      size_t nFg = 0;
      for (size_t i = 0; i < n; ++i) {
        const unsigned v = mask16->GetBufferPointer()[i];
        mask->GetBufferPointer()[i] = v > 0 ? 1 : 0;           // ClampImageFilter(0, 1)
        for (unsigned a : foreground) if (v == a) { fg[i] = 1; break; }
        nFg += fg[i];
      }
      // The features stay on the device; only the sampled voxels' values come back ("compaction
      // sink", ife_cuda_emphysema_feature_samples): all foreground voxels in voxel order (-S 0), or
      // -S random foreground voxels per scale, drawn here exactly as the reference draws them.
      ife::CudaContext& c = ife::CudaContext::Instance();
      const ife::Geometry& g = image->GetGeometry();
      std::vector<double> sig(scales.begin(), scales.end());
      if (nSamples == 0) {
        if (nFg) {
          std::vector<float> rows(scales.size() * numFeatures * nFg);
          size_t got = 0;
          c.Check(ife_cuda_emphysema_feature_samples(c.Handle(), image->GetBufferPointer(), mask->GetBufferPointer(),
                                                     fg.data(), nullptr, 0, g.size.data(), g.spacing.data(), sig.data(),
                                                     (int)sig.size(), 0, rows.data(), &got, IFE_MEM_HOST));
          if (got != nFg) throw std::runtime_error("sample count mismatch");
          for (size_t r = 0; r < scales.size() * numFeatures; ++r)
            samples[r].insert(samples[r].end(), rows.begin() + r * nFg, rows.begin() + (r + 1) * nFg);
        }
      } else {
        if (nFg == 0) throw std::runtime_error("mask has no foreground voxel to sample");
        std::uniform_int_distribution<size_t> pick(0, n - 1);
        for (size_t s = 0; s < scales.size(); ++s) {
          std::vector<long long> index;
          index.reserve(nSamples);
          while (index.size() < nSamples) {
            const size_t i = pick(gen);
            if (fg[i]) index.push_back((long long)i);
          }
          std::vector<float> rows(numFeatures * index.size());
          size_t got = 0;
          c.Check(ife_cuda_emphysema_feature_samples(c.Handle(), image->GetBufferPointer(), mask->GetBufferPointer(),
                                                     nullptr, index.data(), index.size(), g.size.data(), g.spacing.data(),
                                                     &sig[s], 1, 0, rows.data(), &got, IFE_MEM_HOST));
          for (size_t j = 0; j < numFeatures; ++j)
            samples[j + s * numFeatures].insert(samples[j + s * numFeatures].end(), rows.begin() + j * got, rows.begin() + (j + 1) * got);
        }
      }
    } catch (std::exception& e) {
      std::cerr << "Failed to process." << std::endl << "Image: '" << pr.first << "'" << std::endl
                << "Mask: '" << pr.second << "'" << std::endl << "ExceptionObject: " << e.what() << std::endl;
      return EXIT_FAILURE;
    }
  }

  std::ofstream out(outfilePath);
  out << "# Features: GaussianBlur GradientMagnitude Eigenvalue1 Eigenvalue2 Eigenvalue3 LaplacianOfGaussian GaussianCurvature FrobeniusNorm\n"
      << "# Scales: ";
  for (size_t i = 0; i < scales.size(); ++i) out << scales[i] << (i + 1 < scales.size() ? ' ' : '\n');
  if (!out.good()) { std::cerr << "Error writing edges header to file." << std::endl << "Out path: " << outfilePath << std::endl; return EXIT_FAILURE; }
  try {
    ife::CudaContext& c = ife::CudaContext::Instance();
    for (std::vector<float>& row : samples) {
      c.Check(ife_cuda_sort_f32(c.Handle(), row.data(), row.size(), IFE_MEM_HOST));
      std::vector<float> edges;
      ife::determineEdgesForEqualizedHistogram(row.begin(), row.end(), std::back_inserter(edges), nBins);
      for (size_t k = 0; k < edges.size(); ++k) out << (k ? "," : "") << edges[k];
      out << std::endl;
      if (!out.good()) { std::cerr << "Error writing to edges to file." << std::endl << "Out path: " << outfilePath << std::endl; return EXIT_FAILURE; }
    }
  } catch (std::exception& e) {
    std::cerr << "Failed to determine edges: " << e.what() << std::endl;
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}
