// oracle/cvstub: see opencv2/core/core.hpp
#include "opencv2/core/core.hpp"
