// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for <uhd/usrp/multi_usrp.hpp>; UHD is not installed here.
// Only the type names stored by value in /root/reference/include/extensible_cognitive_radio.hpp.
#pragma once
#include <memory>
namespace uhd {
struct time_spec_t { double secs; time_spec_t(double s = 0.0) : secs(s) {} double get_real_secs() const { return secs; } };
struct tx_metadata_t { bool start_of_burst, end_of_burst, has_time_spec; time_spec_t time_spec; };
struct rx_metadata_t { time_spec_t time_spec; int error_code; };
namespace usrp { struct multi_usrp { typedef std::shared_ptr<multi_usrp> sptr; }; }
}
