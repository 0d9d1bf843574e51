// stand-in for <boost/type_traits/type_identity.hpp> (test infrastructure only;
// Boost 1.83.0 is pinned by the reference makefile:27-31 but absent here).
// Used only for the non-deduced-context trick in singlet_CFR.hpp:424-427.
#ifndef B200RT_BOOST_TYPE_IDENTITY_STANDIN
#define B200RT_BOOST_TYPE_IDENTITY_STANDIN
namespace boost { template <class T> struct type_identity { typedef T type; }; }
#endif
