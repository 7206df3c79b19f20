from .run_automoe import build_image_transform, load_model, model_infer  # noqa: F401
