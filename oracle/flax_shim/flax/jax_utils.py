def replicate(x):
    return x
