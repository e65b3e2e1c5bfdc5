// itk:: spellings of the host-side image classes, so that code written against the reference's
// ITK types (itk::Image<float, 3>, itk::VectorImage<float, 3>, ::Pointer, ->GetOutput() ...)
// resolves to the B200 facades.  Only what the reference's hot-path tools use is provided;
// the dimension parameter is accepted and must be 3.
#ifndef IFE_B200_ITK_COMPAT_IMAGE_H
#define IFE_B200_ITK_COMPAT_IMAGE_H
#include "ife/Context.h"
#include "ife/Image.h"

namespace itk {
template <typename TPixel, unsigned int VDimension = 3>
using Image = ife::Image<TPixel>;
template <typename TPixel, unsigned int VDimension = 3>
using VectorImage = ife::VectorImage<TPixel>;
}  // namespace itk
#endif
