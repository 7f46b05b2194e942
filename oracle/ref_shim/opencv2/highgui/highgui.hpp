// TEST INFRASTRUCTURE ONLY.  Stand-in for <opencv2/highgui/highgui.hpp>: nothing on the path uses it.
