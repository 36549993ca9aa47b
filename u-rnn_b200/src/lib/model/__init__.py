from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)   # lets src.lib.model.earlystopping fall through to the reference tree
