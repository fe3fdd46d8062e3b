#include <boost/random.hpp>
