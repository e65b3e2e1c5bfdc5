// itk::ImageFileWriter over the NIfTI-1 writer of ife/IO/NiftiIO.h: SetInput / SetFileName /
// Update (which first updates whatever produces the input).
#ifndef IFE_B200_ITK_COMPAT_IMAGE_FILE_WRITER_H
#define IFE_B200_ITK_COMPAT_IMAGE_FILE_WRITER_H
#include <memory>
#include <string>

#include "ife/IO/NiftiIO.h"
#include "itkImage.h"

namespace itk {
template <typename TInputImage>
class ImageFileWriter {
public:
  typedef ImageFileWriter Self;
  typedef std::shared_ptr<Self> Pointer;
  static Pointer New() { return Pointer(new Self()); }
  void SetInput(const TInputImage* image) { m_Input = image; }
  void SetFileName(const std::string& name) { m_FileName = name; }
  void Update() {
    if (!m_Input) throw ife::ExceptionObject(IFE_E_INVALID, "ImageFileWriter: input not set");
    m_Input->UpdateSource();
    try {
      ife::nifti::Write(m_FileName, m_Input->GetGeometry(), m_Input->GetBufferPointer());
    } catch (const std::exception& e) {
      throw ife::ExceptionObject(IFE_E_INVALID, e.what());
    }
  }
private:
  ImageFileWriter() {}
  const TInputImage* m_Input = nullptr;
  std::string m_FileName;
};
}  // namespace itk
#endif
