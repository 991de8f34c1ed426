// stand-in: data_type.h includes it, the selection sources do not use it
#pragma once
#include <vector>
namespace tbb { template <class T> using concurrent_vector = std::vector<T>; }
