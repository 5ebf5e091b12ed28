// ORACLE shim (test infrastructure): boost::any as /root/reference uses it (include/cpprob/sample.hpp:19-47,
// src/cpprob/state.cpp:183, src/cpprob/sample.cpp:29): a copyable type-erased value with empty() and any_cast<T>.
#ifndef CPPROB_REF_SHIM_BOOST_ANY_HPP
#define CPPROB_REF_SHIM_BOOST_ANY_HPP
#include <memory>
#include <typeinfo>
#include <utility>
namespace boost {
class bad_any_cast : public std::bad_cast {
public:
    const char * what() const noexcept override { return "boost::bad_any_cast (oracle shim)"; }
};
class any {
    struct holder_base {
        virtual ~holder_base() {}
        virtual holder_base * clone() const = 0;
        virtual const std::type_info & type() const = 0;
    };
    template<class T> struct holder : holder_base {
        T v;
        explicit holder(const T & x) : v(x) {}
        holder_base * clone() const override { return new holder(v); }
        const std::type_info & type() const override { return typeid(T); }
    };
    std::unique_ptr<holder_base> p_;
public:
    any() {}
    any(const any & o) : p_(o.p_ ? o.p_->clone() : nullptr) {}
    any(any && o) noexcept : p_(std::move(o.p_)) {}
    template<class T, class = typename std::enable_if<!std::is_same<typename std::decay<T>::type, any>::value>::type>
    any(const T & x) : p_(new holder<typename std::decay<T>::type>(x)) {}
    any & operator=(any o) { p_ = std::move(o.p_); return *this; }
    bool empty() const { return !p_; }
    const std::type_info & type() const { return p_ ? p_->type() : typeid(void); }
    template<class T> friend T any_cast(const any & a);
};
template<class T> T any_cast(const any & a)
{
    typedef typename std::decay<T>::type U;
    if (a.empty() || a.p_->type() != typeid(U)) throw bad_any_cast();
    return static_cast<any::holder<U> *>(a.p_.get())->v;
}
}
#endif
