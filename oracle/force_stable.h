/* Force-included (-include) when building oracle/_ref/fastq-dupaway-stable from the UNMODIFIED reference
 * sources: every std header the sources use is pulled in first, then `sort` is renamed to `stable_sort`
 * so that src/external_sort.hpp:105 and src/paired_external_sort.hpp:135 become std::stable_sort
 * (the reference's choice of representative inside a tie group is an introsort artefact;
 * the stable build defines "first in input order survives" for single-chunk runs).  The member functions
 * ExternalSorter::sort / PairedExternalSorter::sort are renamed consistently, which is harmless. */
#include <algorithm>
#include <any>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <functional>
#include <iostream>
#include <iterator>
#include <list>
#include <map>
#include <memory>
#include <queue>
#include <random>
#include <sstream>
#include <string>
#include <unordered_set>
#include <vector>
#include <zlib.h>
#include <unistd.h>
#define sort stable_sort
