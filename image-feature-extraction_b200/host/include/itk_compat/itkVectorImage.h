#include "itkImage.h"
