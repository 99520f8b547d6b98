// TEST INFRASTRUCTURE ONLY -- see ../opencv.hpp
#include "../opencv.hpp"
