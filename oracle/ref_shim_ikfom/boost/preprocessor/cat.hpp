#pragma once
#include <boost/preprocessor/seq.hpp>
