"""vmatwpencode_b200 - B200 (sm_100a) implementation of the dose / objective / gradient
evaluation and aperture pricing of wilmerhenao/VMATwPenCode.

    synth        synthetic planning cases (the reference ships no data)
    plan         DosePlan: D resident in HBM, evaluation and pricing through the C ABI
    VMATlibrary  the reference's function names (calcObjGrad, PricingProblem, ...) on that plan
    kmedoids     greedy k-medoid ordering with the distance pass on the GPU
    dist         voxel-slab sharding across the GPUs of one node
"""
__version__ = "0.1.0"
