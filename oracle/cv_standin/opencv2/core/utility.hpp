// stand-in: see opencv2/opencv.hpp
#include "../opencv.hpp"
