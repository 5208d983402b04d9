/* TEST INFRASTRUCTURE: see lua.h in this directory. */
#include "lua.h"
