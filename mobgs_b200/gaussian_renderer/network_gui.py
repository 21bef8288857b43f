"""Placeholder for the reference's SIBR viewer socket (gaussian_renderer/network_gui.py), which
train.py imports next to `render` and initialises at start-up (train.py:18, :989).  The viewer
protocol is out of scope: connections are never accepted."""
conn = None
addr = None
host = "127.0.0.1"
port = 6009


def init(wish_host, wish_port):
    global host, port
    host, port = wish_host, wish_port


def try_connect():
    return None
