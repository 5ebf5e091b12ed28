// cpprob-b200: the text grammar of CPProb's posterior files and observation strings.
//
// Same grammar as /root/reference include/cpprob/serialization.hpp — output :41-46 pair "(a b)",
// :58-69 tuple "(a b c)", :71-98 vector / array "[a b c]", :100-105 map "{(k v) (k v)}"; input
// :107-257; parse_file / parse_string :265-284 — written independently as one recursive `io<T>`
// trait instead of a family of overloaded stream operators, so that it cannot hijack operator<< /
// operator>> of unrelated user types.  Used by StatsPrinter (record files) and by the CLI
// (`-o "3 4"`, `-o "[1.5 2 3]"`).
#ifndef INCLUDE_SERIALIZATION_HPP
#define INCLUDE_SERIALIZATION_HPP

#include <array>
#include <cstddef>
#include <fstream>
#include <iostream>
#include <istream>
#include <map>
#include <ostream>
#include <sstream>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

namespace cpprob {
namespace text {

template<class T, class Enable = void>
struct io {   // scalars and anything else with stream operators
    static void write(std::ostream & os, const T & v) { os << v; }
    static bool read(std::istream & is, T & v) { return static_cast<bool>(is >> v); }
};

inline bool expect(std::istream & is, char want)
{
    char ch;
    if (!(is >> std::ws >> ch)) return false;
    if (ch != want) {
        is.putback(ch);
        is.setstate(std::ios_base::failbit);
        return false;
    }
    return true;
}

template<class It>
void write_range(std::ostream & os, It first, It last, char open, char close)
{
    using value_t = typename std::decay<decltype(*first)>::type;
    os << open;
    for (bool head = true; first != last; ++first, head = false) {
        if (!head) os << ' ';
        io<value_t>::write(os, *first);
    }
    os << close;
}

template<class A, class B>
struct io<std::pair<A, B>> {
    static void write(std::ostream & os, const std::pair<A, B> & p)
    {
        os << '(';
        io<A>::write(os, p.first);
        os << ' ';
        io<B>::write(os, p.second);
        os << ')';
    }
    static bool read(std::istream & is, std::pair<A, B> & p)
    {
        typename std::remove_const<A>::type a{};
        if (!expect(is, '(') || !io<typename std::remove_const<A>::type>::read(is, a) || !io<B>::read(is, p.second) || !expect(is, ')')) return false;
        const_cast<typename std::remove_const<A>::type &>(p.first) = std::move(a);
        return true;
    }
};

template<class... T>
struct io<std::tuple<T...>> {
    template<std::size_t I>
    static void write_from(std::ostream &, const std::tuple<T...> &, std::integral_constant<std::size_t, sizeof...(T)>) {}
    template<std::size_t I, std::size_t N>
    static void write_from(std::ostream & os, const std::tuple<T...> & t, std::integral_constant<std::size_t, N>)
    {
        if (I) os << ' ';
        io<typename std::tuple_element<I, std::tuple<T...>>::type>::write(os, std::get<I>(t));
        write_from<I + 1>(os, t, std::integral_constant<std::size_t, I + 1 == sizeof...(T) ? sizeof...(T) : I + 1>());
    }
    static void write(std::ostream & os, const std::tuple<T...> & t)
    {
        os << '(';
        write_from<0>(os, t, std::integral_constant<std::size_t, 0>());
        os << ')';
    }
    template<std::size_t I>
    static bool read_from(std::istream &, std::tuple<T...> &, std::integral_constant<std::size_t, sizeof...(T)>) { return true; }
    template<std::size_t I, std::size_t N>
    static bool read_from(std::istream & is, std::tuple<T...> & t, std::integral_constant<std::size_t, N>)
    {
        if (!io<typename std::tuple_element<I, std::tuple<T...>>::type>::read(is, std::get<I>(t))) return false;
        return read_from<I + 1>(is, t, std::integral_constant<std::size_t, I + 1 == sizeof...(T) ? sizeof...(T) : I + 1>());
    }
    static bool read(std::istream & is, std::tuple<T...> & t)
    {
        return expect(is, '(') && read_from<0>(is, t, std::integral_constant<std::size_t, 0>()) && expect(is, ')');
    }
};
template<>
struct io<std::tuple<>> {
    static void write(std::ostream & os, const std::tuple<> &) { os << "()"; }
    static bool read(std::istream & is, std::tuple<> &) { return expect(is, '(') && expect(is, ')'); }
};

template<class T>
struct io<std::vector<T>> {
    static void write(std::ostream & os, const std::vector<T> & v) { write_range(os, v.begin(), v.end(), '[', ']'); }
    static bool read(std::istream & is, std::vector<T> & v)
    {
        if (!expect(is, '[')) return false;
        for (;;) {
            T item{};
            if (!io<T>::read(is, item)) break;    // the element that fails to parse ends the list ...
            v.emplace_back(std::move(item));
        }
        is.clear();                               // ... and must be the closing bracket (serialization.hpp:157-169)
        return expect(is, ']');
    }
};

template<class T, std::size_t N>
struct io<std::array<T, N>> {
    static void write(std::ostream & os, const std::array<T, N> & v) { write_range(os, v.begin(), v.end(), '[', ']'); }
    static bool read(std::istream & is, std::array<T, N> & v)
    {
        if (!expect(is, '[')) return false;
        for (auto & item : v) {
            if (!io<T>::read(is, item)) return false;
        }
        return expect(is, ']');
    }
};

template<class K, class V>
struct io<std::map<K, V>> {
    static void write(std::ostream & os, const std::map<K, V> & m) { write_range(os, m.begin(), m.end(), '{', '}'); }
    static bool read(std::istream & is, std::map<K, V> & m)
    {
        if (!expect(is, '{')) return false;
        for (;;) {
            std::pair<K, V> item{};
            if (!io<std::pair<K, V>>::read(is, item)) break;
            m.insert(std::move(item));
        }
        is.clear();
        return expect(is, '}');
    }
};

template<class T>
std::string to_string(const T & v)
{
    std::ostringstream os;
    io<T>::write(os, v);
    return os.str();
}

}  // namespace text

// Reads the elements of `tup` one after the other from the text (no enclosing parentheses), as the
// CLI does with the observation string (serialization.hpp:259-284).
namespace detail {
template<std::size_t I, class... T>
typename std::enable_if<I == sizeof...(T), bool>::type parse_elements(std::istream &, std::tuple<T...> &) { return true; }
template<std::size_t I, class... T>
typename std::enable_if<(I < sizeof...(T)), bool>::type parse_elements(std::istream & is, std::tuple<T...> & tup)
{
    using elem_t = typename std::tuple_element<I, std::tuple<T...>>::type;
    if (!text::io<elem_t>::read(is, std::get<I>(tup))) return false;
    return parse_elements<I + 1>(is, tup);
}
}  // namespace detail

template<class... T>
bool parse_string(const std::string & param, std::tuple<T...> & tup)
{
    std::istringstream iss(param);
    return detail::parse_elements<0>(iss, tup);
}

template<class... T>
bool parse_file(const std::string & path, std::tuple<T...> & tup)
{
    std::ifstream file(path);
    if (!file) {
        std::cerr << "File " << path << " could not be opened.\n";
        return false;
    }
    return detail::parse_elements<0>(file, tup);
}

}  // end namespace cpprob
#endif  // INCLUDE_SERIALIZATION_HPP
