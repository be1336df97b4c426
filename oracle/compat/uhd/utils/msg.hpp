// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for <uhd/utils/msg.hpp>; UHD is not installed here.
#pragma once
#include <string>
namespace uhd { namespace msg { enum type_t { status = 's', warning = 'w', error = 'e', fastpath = 'f' }; } }
