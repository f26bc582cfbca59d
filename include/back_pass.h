/* shim: the reference header of this name, provided by ilqg_compat.h */
#include "ilqg_compat.h"
