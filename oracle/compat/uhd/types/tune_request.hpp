// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for <uhd/types/tune_request.hpp> (nothing needed).
#pragma once
