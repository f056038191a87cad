// test stub: see tests/tf_stub/tf_stub.h
#include "../../../tf_stub.h"
