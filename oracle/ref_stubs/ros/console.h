// STAND-IN for <ros/console.h>: the logging macros live in ros/ros.h of this directory.
#pragma once
#include "ros.h"
