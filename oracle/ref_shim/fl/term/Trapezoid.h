#pragma once
#include <fl/Headers.h>
