// ORACLE shim: see path.hpp
#include <boost/filesystem/path.hpp>
