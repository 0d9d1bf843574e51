// declaration-only stand-in for boost's cardinal_cubic_b_spline (test
// infrastructure only).  atm/atmosphere_average_1d.hpp names the type as a data
// member; the oracle never constructs an atmosphere_average_1d, it only needs the
// class to be complete for the dynamic_cast in
// grid_spherical_azimuthally_symmetric.hpp:270.
#ifndef B200RT_BOOST_BSPLINE_STANDIN
#define B200RT_BOOST_BSPLINE_STANDIN
namespace boost { namespace math { namespace interpolators {
template <class T> class cardinal_cubic_b_spline {
public:
  cardinal_cubic_b_spline() {}
  T operator()(T) const { return T(0); }
  T prime(T) const { return T(0); }
};
}}}
#endif
