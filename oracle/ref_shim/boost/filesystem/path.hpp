// ORACLE shim (test infrastructure): the two Boost.Filesystem names StatsPrinter uses
// (/root/reference include/cpprob/postprocess/stats_printer.hpp:27-29): path(string) and exists(path).
#ifndef CPPROB_REF_SHIM_FS_PATH_HPP
#define CPPROB_REF_SHIM_FS_PATH_HPP
#include <string>
#include <sys/stat.h>
namespace boost { namespace filesystem {
class path {
public:
    path() {}
    path(const std::string & s) : s_(s) {}
    path(const char * s) : s_(s) {}
    const std::string & string() const { return s_; }
    const char * c_str() const { return s_.c_str(); }
private:
    std::string s_;
};
inline bool exists(const path & p) { struct stat st; return ::stat(p.c_str(), &st) == 0; }
}}
#endif
