"""Empty stand-in: the reference imports pyglet at module scope of its
visualisation file, which quadrotor_multi.py:18 pulls in; rendering is out of scope."""
