#include "fmt/core.h"
