/* TEST INFRASTRUCTURE ONLY (oracle build shim).
 * mmap-only stand-in for api/tests/halApiTestSupport.cpp (which includes <H5Cpp.h>): implements the
 * members declared in api/tests/halApiTestSupport.h.  Extra behaviour: when the environment variable
 * ORACLE_KEEP_FIXTURE names a directory, the HAL file each test builds is kept there as
 * <mangled-test-class>.hal instead of being unlinked, so the reference's hand-built fixtures can be
 * replayed against the GPU path. */
#include "halApiTestSupport.h"
#include "halAlignmentInstance.h"
#include <cstdlib>
#include <iostream>
#include <typeinfo>
#include <unistd.h>

using namespace std;
using namespace hal;

AlignmentPtr getTestAlignmentInstances(const std::string &storageFormat, const std::string &path, unsigned mode) {
    if (storageFormat != STORAGE_FORMAT_MMAP) {
        throw hal_exception("oracle build is mmap-only");
    }
    return AlignmentPtr(mmapAlignmentInstance(path, mode, 1024 * 1024 * 1024));
}

int runHalTestSuite(int argc, char *argv[], CuSuite *suite) {
    CuString *output = CuStringNew();
    CuSuiteRun(suite);
    CuSuiteSummary(suite, output);
    CuSuiteDetails(suite, output);
    cerr << argv[0] << " " << output->buffer << endl;
    return (suite->failCount > 0) ? 1 : 0;
}

string AlignmentTest::randomString(hal_size_t length) {
    static const char alphabet[] = "acgtACGTNn";
    string s(length, '?');
    for (hal_size_t i = 0; i < length; ++i) {
        s[i] = alphabet[rand() % 10];
    }
    return s;
}

void AlignmentTest::check(CuTest *testCase) {
    _testCase = testCase;
    try {
        checkOne(testCase, STORAGE_FORMAT_MMAP);
    } catch (const exception &e) {
        CuFail(testCase, stString_print("Caught exception while testing: %s", e.what()));
    }
}

void AlignmentTest::checkOne(CuTest *testCase, const string &storageFormat) {
    const char *keep = getenv("ORACLE_KEEP_FIXTURE");
    string path = keep ? string(keep) + "/" + typeid(*this).name() + ".hal" : string(getTempFile());
    {
        AlignmentPtr calignment(getTestAlignmentInstances(storageFormat, path, CREATE_ACCESS));
        _createPath = path;
        createCallBack(calignment);
        calignment->close();
    }
    {
        AlignmentPtr ralignment(getTestAlignmentInstances(storageFormat, path, READ_ACCESS));
        _checkPath = path;
        checkCallBack(ralignment);
        ralignment->close();
    }
    if (!keep) {
        ::unlink(path.c_str());
    }
}
