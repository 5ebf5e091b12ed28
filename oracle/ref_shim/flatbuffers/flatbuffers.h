// ORACLE shim (test infrastructure): the part of the FlatBuffers RUNTIME (third party, absent from this image) that the
// reference's generated protocol header (/root/reference include/flatbuffers/infcomp_generated.h, used unmodified) and
// its CSIS code name.  Everything here only has to COMPILE and link: the wire protocol belongs to the compile / CSIS
// modes, which the SIS path never enters (cpprob.hpp:79-106 branch on State::sis() first).  The builder really does
// nothing; readers return defaults.
#ifndef CPPROB_REF_SHIM_FLATBUFFERS_H
#define CPPROB_REF_SHIM_FLATBUFFERS_H
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>
#define FLATBUFFERS_FINAL_CLASS final
namespace flatbuffers {
typedef uint32_t uoffset_t;
typedef uint16_t voffset_t;
template<class T> struct Offset {
    uoffset_t o;
    Offset() : o(0) {}
    Offset(uoffset_t v) : o(v) {}
    Offset<void> Union() const { return Offset<void>(o); }
};
template<class T> class Vector {
public:
    uoffset_t size() const { return 0; }
    const T * begin() const { return nullptr; }
    const T * end() const { return nullptr; }
    T Get(uoffset_t) const { return T(); }
    template<class E> E GetEnum(uoffset_t) const { return E(); }
    T operator[](uoffset_t) const { return T(); }
    const T * data() const { return nullptr; }
};
template<class T> class Vector<Offset<T>> {          // a vector of tables hands out pointers
public:
    uoffset_t size() const { return 0; }
    const T * Get(uoffset_t) const { return nullptr; }
    const T * operator[](uoffset_t) const { return nullptr; }
    struct iterator {                                 // dereferences to a table pointer, as flatbuffers::VectorIterator does
        const T * operator*() const { return nullptr; }
        const T * operator->() const { return nullptr; }
        iterator & operator++() { return *this; }
        bool operator!=(const iterator &) const { return false; }
        bool operator==(const iterator &) const { return true; }
    };
    iterator begin() const { return iterator(); }
    iterator end() const { return iterator(); }
};
struct String : public Vector<char> {
    const char * c_str() const { return ""; }
    std::string str() const { return std::string(); }
};
class Verifier {
public:
    Verifier(const uint8_t *, size_t) {}
    template<class T> bool VerifyBuffer(const char *) { return true; }
    template<class T> bool VerifyTable(const T *) { return true; }
    template<class T> bool Verify(const T *) const { return true; }
    bool Verify(const void *, size_t) const { return true; }
    template<class T> bool VerifyVectorOfTables(const Vector<Offset<T>> *) { return true; }
    bool EndTable() { return true; }
};
class Table {
public:
    template<class T> T GetField(voffset_t, T defaultval) const { return defaultval; }
    template<class P> P GetPointer(voffset_t) const { return nullptr; }
    bool VerifyTableStart(Verifier &) const { return true; }
    template<class T> bool VerifyField(const Verifier &, voffset_t) const { return true; }
    bool VerifyOffset(const Verifier &, voffset_t) const { return true; }
    bool VerifyOffsetRequired(const Verifier &, voffset_t) const { return true; }
};
class FlatBufferBuilder {
public:
    explicit FlatBufferBuilder(size_t = 1024) {}
    void Clear() {}
    uoffset_t GetSize() const { return 0; }
    uint8_t * GetBufferPointer() const { return nullptr; }
    uoffset_t StartTable() { return 0; }
    uoffset_t EndTable(uoffset_t, voffset_t) { return 0; }
    template<class T> void AddElement(voffset_t, T, T) {}
    template<class T> void AddOffset(voffset_t, Offset<T>) {}
    template<class T> void Required(Offset<T>, voffset_t) {}
    template<class T> Offset<Vector<T>> CreateVector(const std::vector<T> &) { return Offset<Vector<T>>(); }
    template<class T> Offset<Vector<T>> CreateVector(const T *, size_t) { return Offset<Vector<T>>(); }
    Offset<String> CreateString(const std::string &) { return Offset<String>(); }
    Offset<String> CreateString(const char *) { return Offset<String>(); }
    template<class T> void Finish(Offset<T>, const char * = nullptr) {}
};
template<class T> const T * GetRoot(const void * buf) { return static_cast<const T *>(buf); }
}  // namespace flatbuffers
#endif
