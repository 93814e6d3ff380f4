// Inert stand-in for Boost.Python: only what simulator.h / simulator_entity.h need to PARSE.
// The _ref driver builds Entity structs field by field and never touches Python objects.
#pragma once
#include <string>
namespace boost { namespace python {
struct object {
    object() {}
    template <typename T> object(const T&) {}
    template <typename T> object operator[](const T&) const { return object(); }
    template <typename T> object& operator=(const T&) { return *this; }
    object attr(const char*) const { return object(); }
    template <typename... A> object operator()(A...) const { return object(); }
};
struct dict : object { using object::operator=; using object::operator[]; };
struct tuple : object {};
struct list : object { template <typename T> void append(const T&) {} };
template <typename T> struct extract {
    template <typename U> extract(const U&) {}
    operator T() const { return T(); }
};
template <typename... A> tuple make_tuple(A...) { return tuple(); }
inline int len(const object&) { return 0; }
}}  // namespace boost::python
