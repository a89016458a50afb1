// empty stand-in: the compiled factor / feature-manager sources use nothing of OpenCV (parameters.h merely includes it);
// the standard headers below are the ones the real OpenCV headers bring in and the reference relies on transitively
#include <map>
#include <set>
#include <string>
#include <vector>
