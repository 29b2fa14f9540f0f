#include "b2r_mini_glm.hpp"
