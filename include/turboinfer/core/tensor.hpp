// turboinfer/core/tensor.hpp -- host-side value types of the B200 build: DataType, TensorShape, Tensor.
// Same names, members and behaviour as the reference's include/turboinfer/core/tensor.hpp:23-286 (owning,
// zero-initialised host buffer; copies and reshape/slice/clone are deep; data_ptr<T> checks sizeof(T) only).
// Device memory never lives inside a Tensor: it is owned by handles behind the C ABI (include/ti_b200.h).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <initializer_list>
#include <memory>
#include <stdexcept>
#include <vector>

namespace turboinfer {
namespace core {

enum class DataType { kFloat32, kFloat16, kInt32, kInt16, kInt8, kUInt8 };

size_t get_dtype_size(DataType dtype);
const char* dtype_to_string(DataType dtype);

class TensorShape {
public:
    explicit TensorShape(std::initializer_list<size_t> dimensions) : dimensions_(dimensions) { calculate_total_size(); }
    explicit TensorShape(const std::vector<size_t>& dimensions) : dimensions_(dimensions) { calculate_total_size(); }
    size_t ndim() const noexcept { return dimensions_.size(); }
    size_t size(size_t dim) const {
        if (dim >= dimensions_.size()) throw std::out_of_range("Dimension index out of range");
        return dimensions_[dim];
    }
    size_t total_size() const noexcept { return total_size_; }
    const std::vector<size_t>& dimensions() const noexcept { return dimensions_; }
    bool operator==(const TensorShape& o) const noexcept { return dimensions_ == o.dimensions_; }
    bool operator!=(const TensorShape& o) const noexcept { return !(*this == o); }
    bool is_broadcastable_with(const TensorShape& o) const noexcept;

private:
    void calculate_total_size() {
        total_size_ = dimensions_.empty() ? 0 : 1;
        for (size_t d : dimensions_) total_size_ *= d;
    }
    std::vector<size_t> dimensions_;
    size_t total_size_ = 0;
};

class Tensor {
public:
    Tensor(const TensorShape& shape, DataType dtype = DataType::kFloat32);
    Tensor(const TensorShape& shape, const void* data, DataType dtype = DataType::kFloat32);
    Tensor(const Tensor& other);
    Tensor(Tensor&& other) noexcept = default;
    Tensor& operator=(const Tensor& other);
    Tensor& operator=(Tensor&& other) noexcept = default;
    ~Tensor() = default;

    const TensorShape& shape() const noexcept { return shape_; }
    DataType dtype() const noexcept { return dtype_; }
    size_t element_size() const noexcept { return get_dtype_size(dtype_); }
    size_t byte_size() const noexcept { return shape_.total_size() * element_size(); }
    void* data() noexcept { return data_.get(); }
    const void* data() const noexcept { return data_.get(); }
    template <typename T> T* data_ptr() { validate_type<T>(); return reinterpret_cast<T*>(data_.get()); }
    template <typename T> const T* data_ptr() const { validate_type<T>(); return reinterpret_cast<const T*>(data_.get()); }
    bool empty() const noexcept { return !data_ || shape_.total_size() == 0; }
    template <typename T> void fill(T value) {
        T* p = data_ptr<T>();
        for (size_t i = 0; i < shape_.total_size(); ++i) p[i] = value;
    }
    Tensor clone() const { return Tensor(*this); }
    Tensor reshape(const TensorShape& new_shape) const;
    Tensor slice(const std::vector<size_t>& start, const std::vector<size_t>& end) const;

private:
    template <typename T> void validate_type() const {
        if (sizeof(T) != element_size()) throw std::runtime_error("Type size mismatch with tensor data type");
    }
    TensorShape shape_;
    DataType dtype_;
    std::unique_ptr<uint8_t[]> data_;
};

}  // namespace core
}  // namespace turboinfer
