// Empty: the reference TUs compiled into oracle/_ref include this header but use nothing from it.
#pragma once
#include "../glm.hpp"
