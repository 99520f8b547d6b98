// TEST INFRASTRUCTURE ONLY
#include "se3.hpp"
