from .model import import_model, save_model, copy_state_dict, MobileURSONetB200  # noqa: F401
