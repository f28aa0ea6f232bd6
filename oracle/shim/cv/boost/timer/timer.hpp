// oracle/shim/cv/boost/timer/timer.hpp -- TEST INFRASTRUCTURE ONLY: the reference's headers include it, only its
// drivers (src/cvo_main.cpp:32,48) use it.
#ifndef CVO_ORACLE_SHIM_BOOST_TIMER_HPP
#define CVO_ORACLE_SHIM_BOOST_TIMER_HPP
#include <string>
namespace boost { namespace timer {
class cpu_timer {
  public:
    std::string format() const { return std::string(); }
};
}}  // namespace boost::timer
#endif
