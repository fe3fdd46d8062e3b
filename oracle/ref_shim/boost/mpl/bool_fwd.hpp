#include <boost/mpl/bool.hpp>
