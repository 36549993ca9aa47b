"""B200-native drop-in for the reference package src.lib.model.networks (ConvRNN, encoder, decoder, model,
net_params, utils, head).  Modules not provided here (e.g. losses.py) resolve to any other
src/lib/model/networks directory further down sys.path (the unmodified reference tree)."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
