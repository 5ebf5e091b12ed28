// ORACLE shim: see uuid.hpp
#include <boost/uuid/uuid.hpp>
