// TEST INFRASTRUCTURE — see ../toml.hpp (src/colour.hpp:4 includes this path).
#include "../toml.hpp"
