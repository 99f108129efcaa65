"""gpflow.models.GPModel / InternalDataTrainingLossMixin stand-ins (TEST ONLY)."""


class GPModel:
    def __init__(self, kernel, likelihood, mean_function=None, num_latent_gps=None):
        self.kernel = kernel
        self.likelihood = likelihood
        self.mean_function = mean_function
        self.num_latent_gps = num_latent_gps

    @staticmethod
    def calc_num_latent_gps_from_data(data, kernel, likelihood):
        return data[1].shape[-1]


class InternalDataTrainingLossMixin:
    def training_loss(self):
        return -self.maximum_log_likelihood_objective()
