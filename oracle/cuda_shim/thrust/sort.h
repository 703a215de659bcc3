#include "device_vector.h"
