// empty stand-in: the compiled factor sources use nothing of OpenCV (parameters.h merely includes it)
