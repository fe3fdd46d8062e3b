#include "logical.hpp"
