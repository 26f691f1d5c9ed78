def stop_gradient(x):
    return x
def complex(re, im):
    import numpy as np
    return (np.asarray(re) + 1j * np.asarray(im)).astype(np.complex64)
