// placeholder
template <typename T> int Sampler<T>::get_map(int, double*, double*, double*, int*) { return fail("get_map: not built yet"); }
