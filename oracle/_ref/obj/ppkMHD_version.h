#pragma once
#include <iostream>
inline void print_version_info() { std::cout << "ppkMHD reference (oracle/_ref build, plain Makefile)\\n"; }
