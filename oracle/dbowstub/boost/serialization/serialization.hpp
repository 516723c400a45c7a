// stand-in for <boost/serialization/serialization.hpp>: the vendored DBoW2 only names these in its (unused here) serialize() members.
// TEST INFRASTRUCTURE ONLY (oracle/_ref/libref_bow.so).
#pragma once
namespace boost { namespace serialization {
class access {};
template <class Base, class Derived> Base& base_object(Derived& d) { return static_cast<Base&>(d); }
}}
