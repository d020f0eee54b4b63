"""Hook between the cross-validation drivers and the device classes."""


def prepare_summary(model, predict_config):
    ''' Gibbs / ICM models of this package can average their draws on the GPU while they run
        (summarise_on_device) when burn_in and thinning are known before train(): the drivers know
        them from predict_config, so a fold never stores or copies its all_U / all_V traces. '''
    if hasattr(model, 'summarise_on_device') and predict_config and \
            'burn_in' in predict_config and 'thinning' in predict_config:
        model.summarise_on_device(predict_config['burn_in'], predict_config['thinning'])
