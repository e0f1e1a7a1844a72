#include "pybind11.h"
