// ORACLE shim: see components.hpp
#include <boost/function_types/components.hpp>
