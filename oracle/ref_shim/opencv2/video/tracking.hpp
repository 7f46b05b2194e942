// TEST INFRASTRUCTURE: empty stand-in (flow_reader.cpp includes it; nothing of it is used on this path).
#pragma once
