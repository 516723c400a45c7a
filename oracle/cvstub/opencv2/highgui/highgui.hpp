#include "../cvstub.hpp"
