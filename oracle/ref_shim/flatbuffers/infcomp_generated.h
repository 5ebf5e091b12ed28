// ORACLE shim (test infrastructure).  The reference's log-pdf headers (include/cpprob/distributions/utils_*.hpp) define,
// next to `logpdf<D>`, the CSIS wire-format traits of the same distribution, whose member templates name FlatBuffers
// types.  Those members are never instantiated on the SIS path; this stand-in (found before the reference's generated
// header, which needs the absent flatbuffers/flatbuffers.h) only declares the names they mention.
#ifndef CPPROB_REF_SHIM_INFCOMP_GENERATED_H
#define CPPROB_REF_SHIM_INFCOMP_GENERATED_H
namespace flatbuffers {
template<class T> struct Offset { Offset<void> Union() const; };
class FlatBufferBuilder {
public:
    template<class T, class V> Offset<void> CreateVector(const V &);
};
}
namespace protocol {
struct Discrete;
struct Poisson;
struct UniformDiscrete;
struct Normal;
struct UniformContinuous;
struct MultivariateNormal;
struct NDArray;
template<class... A> flatbuffers::Offset<Discrete> CreateDiscrete(A &&...);
template<class... A> flatbuffers::Offset<Poisson> CreatePoisson(A &&...);
template<class... A> flatbuffers::Offset<UniformDiscrete> CreateUniformDiscrete(A &&...);
template<class... A> flatbuffers::Offset<Normal> CreateNormal(A &&...);
template<class... A> flatbuffers::Offset<UniformContinuous> CreateUniformContinuous(A &&...);
template<class... A> flatbuffers::Offset<MultivariateNormal> CreateMultivariateNormal(A &&...);
template<class... A> flatbuffers::Offset<NDArray> CreateNDArray(A &&...);
}
#endif
