// oracle/ref_build/cooling_stubs.cpp -- TEST INFRASTRUCTURE (not reference source).
// HDF5 is absent from this image.  The reference's cooling-table readers
// (src/cooling/{Grackle,Cloudy}DataReader.cpp) are the only HDF5 users and are reached only when
// cooling.enabled=1, which none of the hot-path configs sets.  These link-time stand-ins abort if
// ever called.
#include "AMReX.H"
#include "cooling/CloudyDataReader.hpp"
#include "cooling/GrackleDataReader.hpp"

namespace quokka::GrackleLikeCooling
{
void initialize_cloudy_data(grackle_data & /*d*/, char const * /*group*/, std::string & /*file*/, code_units & /*u*/)
{
	amrex::Abort("oracle/_ref build has no HDF5: cooling tables unavailable");
}
auto extract_2d_table(amrex::Table3D<double> const & /*t*/, int /*z*/) -> amrex::TableData<double, 2>
{
	amrex::Abort("oracle/_ref build has no HDF5");
	return {};
}
auto copy_1d_table(amrex::Table1D<double> const & /*t*/) -> amrex::TableData<double, 1>
{
	amrex::Abort("oracle/_ref build has no HDF5");
	return {};
}
} // namespace quokka::GrackleLikeCooling

namespace quokka::TabulatedCooling
{
void initialize_cloudy_data(cloudy_cooling_tools_data & /*d*/, std::string const & /*file*/, code_units const & /*u*/)
{
	amrex::Abort("oracle/_ref build has no HDF5: cooling tables unavailable");
}
auto extract_2d_table(amrex::Table2D<double> const & /*t*/) -> amrex::TableData<double, 2>
{
	amrex::Abort("oracle/_ref build has no HDF5");
	return {};
}
auto copy_1d_table(amrex::Table1D<double> const & /*t*/) -> amrex::TableData<double, 1>
{
	amrex::Abort("oracle/_ref build has no HDF5");
	return {};
}
} // namespace quokka::TabulatedCooling
