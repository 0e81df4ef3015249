// ref_gtests_main.cpp -- runs the reference's OWN gtest headers (test/tVX_Material.h, tVX_MaterialLink.h, tVX_Voxel.h,
// tVoxelyze.h; test/VoxelyzeUnitTests.cpp:2-8 of the reference) UNMODIFIED.  The headers are not copied: the build
// (voxelyze_b200/build.py build_ref_gtests) compiles them where they lie under /root/reference through a directory of
// symbolic links, whose sibling `include` link points either at the reference's own headers (CPU) or at the façade's
// (B200), so that the tests' `#include "../include/Voxelyze.h"` picks the implementation under test.
// Left out like SURVEY.md section 4 says: tVX_MaterialVoxel.h (includes a path that does not exist and calls a setter
// that does not exist) and tArray3D.h (raw inverted asserts that abort).  Expected on the unmodified reference: 49 of 51
// pass; CVoxelyze.poissonsSmall and CVoxelyze.deformableMaterialPossions carry stale golden values.
#include "gtest/gtest.h"
#include "reftests/test/tVX_Material.h"
#include "reftests/test/tVX_MaterialLink.h"
#include "reftests/test/tVX_Voxel.h"
#include "reftests/test/tVoxelyze.h"

#include <type_traits>
// the tests call abs() on doubles unqualified (tVoxelyze.h:101,557): the headers under test must bring the floating overload
static_assert(std::is_same<decltype(abs(1.5)), double>::value, "abs(double) must not be the integer abs");

int main(int argc, char** argv)
{
    return ::testing::run_all(argc > 1 ? argv[1] : nullptr) > 250 ? 1 : 0;      // the caller reads the per-test lines
}
