"""Drop-in ``core`` package of the B200-native VAE^2 path (see ../_fallthrough.py)."""
# sub-modules this tree lacks (the reference's control plane) resolve in a reference lib/ later on sys.path
__path__ = __import__("pkgutil").extend_path(__path__, __name__)
