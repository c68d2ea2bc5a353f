#include "THCGeneral.h"
