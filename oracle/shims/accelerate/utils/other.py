def extract_model_from_parallel(model, *a, **k):
    return model
