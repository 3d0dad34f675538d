// stand-in for <boost/thread.hpp>
#pragma once
#include "thread/shared_mutex.hpp"
