#include "starpu.h"
