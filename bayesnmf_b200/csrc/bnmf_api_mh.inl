// placeholder
template <typename T> int Sampler<T>::mh_setup() { return 0; }
template <typename T> int Sampler<T>::mh_iteration(int, uint32_t) { return fail("MH / Normal models: not built yet"); }
template <typename T> int Sampler<T>::rank_sweep() { return fail("rank learning: not built yet"); }
template <typename T> int Sampler<T>::refresh_metrics_only() { return fail("init with supplied Z: not built yet"); }
template <typename T> int Sampler<T>::init_sigmasq_prior() { return 0; }
