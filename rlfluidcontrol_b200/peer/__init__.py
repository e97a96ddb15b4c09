"""Protocol-compatible stand-in for the reference agent server (server/server.py), for images without TensorFlow."""
