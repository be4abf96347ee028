// shim: boost::timer::auto_cpu_timer as the reference's drivers use it (start/stop/report around phases; output not needed)
#pragma once
namespace boost { namespace timer { struct auto_cpu_timer { void start() {} void stop() {} void report() {} void resume() {} }; } }
