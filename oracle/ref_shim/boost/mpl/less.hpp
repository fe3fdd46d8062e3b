#include "comparison.hpp"
