// placeholder until the SDIRK restatement lands
#include "dsb_oracle.hpp"
namespace orc { Method* new_sdirk(const Problem&, int, int* err) { *err = ST_BAD_ARG; return nullptr; } }
