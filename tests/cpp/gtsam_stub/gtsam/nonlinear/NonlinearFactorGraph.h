#include <gtsam/base_stub.h>
