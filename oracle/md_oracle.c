#include "md_oracle.h"
