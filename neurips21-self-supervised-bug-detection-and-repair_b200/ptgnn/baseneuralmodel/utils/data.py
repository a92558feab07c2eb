from typing import Callable, Generic, Iterable, Iterator, TypeVar

T = TypeVar("T")


class LazyDataIterable(Generic[T], Iterable[T]):
    """Re-iterable view over a callable that returns a fresh iterator (reference use: train.py:76-91)."""

    def __init__(self, base_iterable_func: Callable[[], Iterator[T]]):
        self.__base_iterable_func = base_iterable_func

    def __iter__(self) -> Iterator[T]:
        return self.__base_iterable_func()


class MemorizedDataIterable(Generic[T], Iterable[T]):
    """Materialises the first pass and replays it afterwards."""

    def __init__(self, base_iterable: Iterable[T]):
        self.__base = base_iterable
        self.__cache = None

    def __iter__(self) -> Iterator[T]:
        if self.__cache is None:
            self.__cache = list(self.__base)
        return iter(self.__cache)
